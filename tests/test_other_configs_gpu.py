"""BASELINE configs 3 and 5 at a size the reference finishes in seconds on the box's host cores: a variation graph with
multi-seed reads and decoy seeds, and 50 kbp reads through tangle-heavy regions over a sweep of band widths.
The CUDA path (through the C ABI) must reproduce the reference bit for bit (score, range, every mapping, trace fingerprint)."""
import os

import pytest

from graphaligner_b200.tools import gacase, synth
from helpers import REF_ALIGN, assert_same, run_reference

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")]


@pytest.fixture(scope="module")
def api():
    from graphaligner_b200 import api as a
    a.load_library()
    return a


def _check(api, tmp_path, case, what):
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=os.cpu_count() or 4)
    aligner = api.Aligner(api.Graph.from_case(case))
    res = aligner.align(case.reads, case.b, case.B)
    got = res.as_dicts()
    assert sum(1 for x in got if not x["failed"]) >= len(got) * 0.9, what
    assert_same(got, expected, what)
    res.free()
    aligner.close()


def test_config3_variation_graph_multiseed(api, tmp_path):
    # configs[2] scaled: 1 Mbp variation graph (SNP / indel bubbles every 100 bp, inversions), 400 reads x 10 kbp,
    # seeds at read offsets 0 / 5000 / end-300 plus a decoy seed each (the seed loop's "already aligned" pruning)
    g, kw = synth.config3(scale=0.01)
    case = synth.make_case(3, g, 400, kw["read_len"], b=kw["b"], seed_offsets=kw["seed_offsets"], decoys=kw["decoys"], errors=(0.05, 0.05, 0.05))
    _check(api, tmp_path, case, "config 3")


@pytest.mark.parametrize("band", [5, 20, 50, 100])
def test_config5_ultralong_reads_band_sweep(api, tmp_path, band):
    # configs[4] scaled: 50 kbp reads (782 slices) over a 1 Mbp graph with a tangle every 100 kbp, band 5 .. 100
    g = synth.make_graph(5, 1_000_000, chop=32, bubble_every=100, indel_frac=0.2, tangle_every=100_000)
    case = synth.make_case(5, g, 16, 50_000, b=band, errors=(0.05, 0.05, 0.05))
    _check(api, tmp_path, case, "config 5, band %d" % band)


@pytest.mark.parametrize("world", [2, 4])
def test_config4_gfa_graph_sharded_reads(api, tmp_path, world):
    # configs[3] scaled: a GFA graph loaded from its file (GfaGraph + BigraphToDigraph semantics), 15 kbp reads (235 slices),
    # the read set sharded the way the multi-GPU driver shards it (contiguous by index, graph replicated: one context and
    # one upload per shard); the concatenation of the shards' results must equal the reference's run over the whole set
    from graphaligner_b200 import multi_gpu
    from graphaligner_b200.tools import vgio
    g, kw = synth.config4(scale=1.0 / 3000)
    case = synth.make_case(4, g, 96, kw["read_len"], b=kw["b"], errors=(0.05, 0.05, 0.05))
    case.gfa_overlap = kw["gfa_overlap"]
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=os.cpu_count() or 4)
    graph_path, _, _ = vgio.write_case_files(case, str(tmp_path / "c4"))
    got = []
    for lo, hi in multi_gpu.shard_bounds(len(case.reads), world):
        aligner = api.Aligner(api.Graph.load(graph_path))
        res = aligner.align(case.reads[lo:hi], case.b, case.B)
        got.extend(res.as_dicts())
        res.free()
        aligner.close()
    assert sum(1 for x in got if not x["failed"]) >= len(got) * 0.9
    assert_same(got, expected, "config 4, %d shards" % world)


@pytest.mark.parametrize("chop", [20_000, 99_000])
def test_long_nodes_below_the_alternate_cutoff(api, tmp_path, chop):
    # unchopped unitig-like nodes: whole-node banding puts 2 nodes = up to 198 000 columns into every slice, just under the
    # reference's 200 000-bp switch to calculateSliceAlternate (GraphAlignerCommon.h:10); still bit-exact
    g = synth.make_graph(31, 700_000, chop=chop)
    case = synth.make_case(31, g, 4, 600, b=10, errors=(0.03, 0.03, 0.03))
    _check(api, tmp_path, case, "long nodes %d" % chop)


def _crossing_case(node_len):
    # two long nodes B -> C and a read made of the last 300 bp of B and the first 300 bp of C, seeded on B: the forward
    # pass walks from B into C, so the band is both nodes
    import numpy as np
    rng = np.random.default_rng(5)
    b_seq, c_seq = ("".join("ACGT"[i] for i in rng.integers(0, 4, node_len)) for _ in range(2))
    return gacase.Case([(2, b_seq), (3, c_seq)], [(2, False, 3, False)], [("cross", b_seq[-300:] + c_seq[:300], [(2, 0, False)])], 10, 0)


def test_band_just_below_the_alternate_cutoff(api, tmp_path):
    # 2 x 99 000 columns in one band: the widest band the bit-parallel path ever sees (GraphAligner.h:2483)
    _check(api, tmp_path, _crossing_case(99_000), "band of 198 000 columns")


def test_alternate_method_band_is_a_per_read_failure(api):
    # 2 x 101 000 columns: the reference switches to calculateSliceAlternate - and at this commit segfaults there in both
    # build flavours (oracle/_ref/ref_align and ref_align_stock, probed on exactly this case), so there is nothing to
    # compare with.  Here the read comes back failed with GA_FLAG_STREAM_ERROR and the context stays usable.
    case = _crossing_case(101_000)
    aligner = api.Aligner(api.Graph.from_case(case))
    res = aligner.align(case.reads, case.b, case.B)
    d = res.as_dicts()
    assert d[0]["failed"] == 1 and d[0]["flags"] & 1
    res.free()
    # (a read inside B does the same: the seed node starts at score 0 everywhere, so B's last column is within reach and C joins
    # the band of the second slice.)  The failure is per read, not per context: the next batch is served as usual
    d = aligner.align(case.reads * 2, case.b, case.B).as_dicts()
    assert [x["failed"] for x in d] == [1, 1] and all(x["flags"] & 1 for x in d)
    aligner.close()


def test_many_reads_on_long_node_graph_are_split_not_dropped(api, tmp_path):
    # 2 kbp unitig-like nodes and a few thousand streams: the per-stream scratch of the general layout grows with the node
    # length (the band holds whole nodes), so the batch planner has to count it - and when the device is (made) too small for
    # the batch, the batch is cut into several launches instead of failing as a whole (GA_MEM_BUDGET_MB forces the cut)
    g = synth.make_graph(33, 400_000, chop=2000, snp_every=500)
    case = synth.make_case(33, g, 1500, 1000, b=10, errors=(0.04, 0.04, 0.04))
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=os.cpu_count() or 4)
    for budget in (None, "600"):
        if budget:
            os.environ["GA_MEM_BUDGET_MB"] = budget
        try:
            aligner = api.Aligner(api.Graph.from_case(case))
            res = aligner.align(case.reads, case.b, case.B)
            got = res.as_dicts()
            assert_same(got, expected, "long-node graph, budget %s" % budget)
            if budget:
                assert aligner.stats()["launches"] > 3   # more than one launch sequence
            res.free()
            aligner.close()
        finally:
            os.environ.pop("GA_MEM_BUDGET_MB", None)
