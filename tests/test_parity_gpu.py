"""Parity tests proper: the CUDA path (through the C ABI) against the reference's outputs.
Bar: bit-exact scores, alignment ranges, mappings and full trace (FNV fingerprint of every trace item)."""
import os

import pytest

from graphaligner_b200.tools import gacase, synth
from helpers import GOLDEN, REF_ALIGN, assert_same, load_expected, run_reference

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from graphaligner_b200 import api as a
    a.load_library()
    return a


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_fixture(api, golden_dir, name):
    case = gacase.read_case(os.path.join(golden_dir, name + ".gacase"))
    expected = load_expected(os.path.join(golden_dir, name + ".expected"))
    graph = api.Graph.from_case(case)
    aligner = api.Aligner(graph)
    res = aligner.align(case.reads, case.b, case.B)
    assert_same(res.as_dicts(), expected, name)
    res.free()
    aligner.close()


def test_smallexample_known_answer(api, golden_dir):
    # the reference's only shipped fixture (test/smallexample): NDEBUG build -> score 25, node 6738+ offset 0,
    # one edit 41/65, alignmentEnd 128, 65 trace items
    case = gacase.read_case(os.path.join(golden_dir, "smallexample.gacase"))
    aligner = api.Aligner(api.Graph.from_case(case))
    d = aligner.align(case.reads, 10, 0).as_dicts()[0]
    assert (d["failed"], d["score"], d["start"], d["end"], d["ntrace"]) == (0, 25, 0, 128, 65)
    assert d["mappings"] == [(6738 * 2, 0, 0, 41, 65)]
    aligner.close()


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,kw,nreads,rl,b", [
    (101, dict(chop=32, snp_every=1000), 64, 5000, 10),
    (102, dict(chop=32, bubble_every=100, inversion_every=3000), 64, 4000, 10),
    (103, dict(chop=16, bubble_every=50, indel_frac=0.5), 48, 2000, 35),
    (104, dict(chop=32, bubble_every=100, tangle_every=8000, tangle_levels=8), 32, 3000, 20),
])
def test_against_reference_run_here(api, tmp_path, seed, kw, nreads, rl, b):
    # seeded inputs, the reference itself run on the box's CPU as the checker
    g = synth.make_graph(seed, 60000, **kw)
    case = synth.make_case(seed, g, nreads, rl, b=b, seed_offsets=(0, rl // 2), decoys=1)
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=4)
    aligner = api.Aligner(api.Graph.from_case(case))
    res = aligner.align(case.reads, case.b, case.B)
    assert_same(res.as_dicts(), expected, "seed %d" % seed)
    aligner.close()


def test_empty_and_degenerate_inputs(api, golden_dir):
    case = gacase.read_case(os.path.join(golden_dir, "dag_snp.gacase"))
    aligner = api.Aligner(api.Graph.from_case(case))
    # empty batch
    assert len(aligner.align([], 10, 0).as_dicts()) == 0
    name, seq, seeds = case.reads[0]
    reads = [
        ("noseeds", seq, []),                                   # "has no seed hits" -> failed
        ("badnode", seq, [(999999, 0, False)]),                  # unknown node: reference throws out_of_range
        ("badpos", seq, [(seeds[0][0], len(seq) + 5, False)]),   # seed beyond the read
        ("badchar", seq[:50] + "X" + seq[51:], seeds),           # the reference aborts on 'X'
        ("ok", seq, seeds),
    ]
    d = aligner.align(reads, 10, 0).as_dicts()
    assert [x["failed"] for x in d] == [1, 1, 1, 1, 0]
    assert d[1]["flags"] & 2 and d[2]["flags"] & 2 and d[3]["flags"] & 4
    aligner.close()


def test_oversized_batch_is_split_into_launches(api, golden_dir, monkeypatch):
    # a batch whose DP history exceeds the device budget is aligned in several launches with identical results
    case = gacase.read_case(os.path.join(golden_dir, "bubbles_multiseed.gacase"))
    expected = load_expected(os.path.join(golden_dir, "bubbles_multiseed.expected"))
    monkeypatch.setenv("GA_MEM_BUDGET_MB", "12")
    aligner = api.Aligner(api.Graph.from_case(case))
    res = aligner.align(case.reads, case.b, case.B)
    assert aligner.stats()["launches"] >= 4   # two kernels per launch, at least two launches
    assert_same(res.as_dicts(), expected, "chunked")
    aligner.close()


def test_ramp_redo_fixture_fires_the_ramp(api, golden_dir):
    # the ramp_redo golden (bit-exact in test_golden above) is only worth something if streams really go back and redo a
    # stretch with the wide band on it: GA_FLAG_RAMP_REDO (16) on most of its reads
    case = gacase.read_case(os.path.join(golden_dir, "ramp_redo.gacase"))
    assert case.B > case.b
    aligner = api.Aligner(api.Graph.from_case(case))
    d = aligner.align(case.reads, case.b, case.B).as_dicts()
    assert sum(1 for x in d if x["flags"] & 16) >= 6
    aligner.close()


@pytest.mark.parametrize("name,reads", [("ramp_stale", ["read_6"]), ("ramp_stale_long", ["read_6", "read_10"])])
def test_ramp_stale_fixtures_take_the_stale_checkpoint_path(api, golden_dir, name, reads):
    # the ramp_stale goldens (bit-exact in test_golden above) pin the reference's stale-checkpoint behaviour only if those reads
    # really go through it here: GA_FLAG_RAMP_STALE (32) = a stretch was re-computed from a checkpoint of the abandoned pass
    case = gacase.read_case(os.path.join(golden_dir, name + ".gacase"))
    aligner = api.Aligner(api.Graph.from_case(case))
    d = aligner.align(case.reads, case.b, case.B).as_dicts()
    stale = [x["name"] for x in d if x["flags"] & 32]
    assert set(reads) <= set(stale), stale
    assert all(x["flags"] & 16 for x in d if x["flags"] & 32)
    aligner.close()


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
def test_ramp_against_reference_run_here(api, tmp_path):
    # -B differential: narrow bands that lose noisy reads, a wide backup band (GraphAligner.h:2612-2719).  The reference itself
    # crashes on about a third of the inputs where a redo fires (it reads slices its stale checkpoint's band does not hold:
    # profiles/r02_ramp_fuzz.txt); every case it survives must be identical, redo or not, stale checkpoint or not.
    from graphaligner_b200.tools import fuzz
    fuzz.RAMP = True
    try:
        cases = [fuzz.make_case(seed)[0] for seed in range(30300, 30340)]
    finally:
        fuzz.RAMP = False
    compared = redo = stale = 0
    for i, case in enumerate(cases):
        path = str(tmp_path / ("ramp%d.gacase" % i))
        gacase.write_case(case, path)
        try:
            expected, _ = run_reference(path)
        except RuntimeError:
            continue   # the reference crashed
        aligner = api.Aligner(api.Graph.from_case(case))
        d = aligner.align(case.reads, case.b, case.B).as_dicts()
        aligner.close()
        assert_same(d, expected, "ramp fuzz %d" % (30300 + i))
        compared += 1
        redo += sum(1 for x in d if x["flags"] & 16)
        stale += sum(1 for x in d if x["flags"] & 32)
    assert compared >= 20 and redo >= 50 and stale >= 1, (compared, redo, stale)


HOSTSIM = os.path.join(os.path.dirname(REF_ALIGN), "libga_hostsim.so")


@pytest.mark.skipif(not os.path.exists(HOSTSIM), reason="oracle/_ref/libga_hostsim.so not built")
def test_ramp_big_batch_matches_cpu_emulation(api, tmp_path):
    # 3 000 noisy reads with -b 2 -B 22: about 900 streams redo a stretch and 300 go through a stale checkpoint.  The reference
    # cannot check a batch like this (one read it crashes on ends its process), so the device - 32 streams per warp in lock
    # step, lanes in different phases of the slice loop - is compared with the CPU emulation of the same code (one stream at a
    # time), which the reference pins case by case (test_ramp_against_reference_run_here, profiles/r02_ramp_fuzz.txt).
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probe = os.path.join(root, "profiles", "tools", "ramp_big_batch_probe.py")
    outs = []
    for lib in (None, HOSTSIM):
        env = dict(os.environ)
        env.pop("GA_LIB", None)
        if lib:
            env["GA_LIB"] = lib
        out = str(tmp_path / ("ramp_%s.txt" % ("emu" if lib else "gpu")))
        r = subprocess.run([sys.executable, probe, out, "3000"], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-800:]
        outs.append(open(out).read())
    assert outs[0] == outs[1]
    lines = outs[0].split("\n")
    assert sum(1 for l in lines if " flags=" in l and int(l.split(" flags=")[1].split()[0]) & 32) >= 100


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
def test_small_band_overflow_reruns_with_general_layout(api, tmp_path, monkeypatch):
    # 32-bp nodes and band 10 select the small-band layout (16 nodes per band in shared memory); tangles hold far more
    # nodes per band, so those streams must report an overflow and be re-run with the general layout - same results
    g = synth.make_graph(105, 60000, chop=32, snp_every=1000, tangle_every=12000, tangle_levels=12, tangle_width=4, tangle_node=2)
    case = synth.make_case(105, g, 48, 3000, b=10, seed_offsets=(0,))
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=4)
    aligner = api.Aligner(api.Graph.from_case(case))
    res = aligner.align(case.reads, case.b, case.B)
    assert aligner.stats()["retries"] > 0
    assert_same(res.as_dicts(), expected, "small-band overflow")
    res.free()
    # and the general layout alone gives the same answer
    monkeypatch.setenv("GA_NO_SMEM", "1")
    res = aligner.align(case.reads, case.b, case.B)
    assert_same(res.as_dicts(), expected, "general layout")
    aligner.close()


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,kw,rl,b", [
    (106, dict(chop=16, bubble_every=120, cycle_every=400), 2000, 10),
    (107, dict(chop=32, snp_every=150, cycle_every=300, inversion_every=900), 3000, 20),
    (108, dict(chop=8, bubble_every=60, cycle_every=250), 1000, 5),
])
def test_cyclic_graphs_against_reference_run_here(api, tmp_path, seed, kw, rl, b):
    # short cycles inside the band: scores and paths depend on how the reference iterates a cyclic component
    # (last-call node minima, partly confirmed columns); the device replays that schedule
    g = synth.make_graph(seed, 40000, **kw)
    case = synth.make_case(seed, g, 40, rl, b=b, seed_offsets=(0, rl // 3))
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=4)
    aligner = api.Aligner(api.Graph.from_case(case))
    res = aligner.align(case.reads, case.b, case.B)
    d = res.as_dicts()
    assert any(x["flags"] & 8 for x in d), "no read met a cyclic band: the case does not test what it says"
    assert_same(d, expected, "seed %d" % seed)
    aligner.close()


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_pipeline_of_batches_matches_golden(api, golden_dir, depth):
    # ga_pipeline_*: a read set cut into small batches and streamed through `depth` contexts comes back in submission
    # order and equals the reference's output for the whole set (batches of uneven size, more batches than lanes)
    case = gacase.read_case(os.path.join(golden_dir, "bubbles_multiseed.gacase"))
    expected = load_expected(os.path.join(golden_dir, "bubbles_multiseed.expected"))
    n = len(case.reads)
    cuts = sorted(set([0, 1, n // 5, n // 2, n // 2 + 1, n - 2, n]))
    batches = [case.reads[lo:hi] for lo, hi in zip(cuts[:-1], cuts[1:])]
    pipe = api.Pipeline(api.Graph.from_case(case), depth=depth)
    got = []
    for res in pipe.align_all(batches, case.b, case.B):
        got.extend(res.as_dicts())
        res.free()
    assert_same(got, expected, "pipeline depth %d" % depth)
    assert pipe.stats()["streams"] > 0
    pipe.close()


def test_pipeline_full_and_empty_are_reported(api, golden_dir):
    case = gacase.read_case(os.path.join(golden_dir, "dag_snp.gacase"))
    pipe = api.Pipeline(api.Graph.from_case(case), depth=2)
    with pytest.raises(RuntimeError, match="nothing in flight"):
        pipe.next()
    packed = api.PackedReads(case.reads, case.b, case.B)
    pipe.submit(packed)
    pipe.submit(packed)
    with pytest.raises(RuntimeError, match="pipeline full"):
        pipe.submit(packed)
    a = pipe.next().as_dicts()
    b = pipe.next().as_dicts()
    assert [x["score"] for x in a] == [x["score"] for x in b]
    assert pipe.in_flight() == 0
    pipe.close()


def test_pipeline_degenerate_batches_and_split(api, golden_dir, monkeypatch):
    # empty batches, failing reads and batches that each lane must split into several launches (half of a 12 MB budget
    # per lane), mixed in one stream: order and results as for single calls
    case = gacase.read_case(os.path.join(golden_dir, "bubbles_multiseed.gacase"))
    expected = load_expected(os.path.join(golden_dir, "bubbles_multiseed.expected"))
    monkeypatch.setenv("GA_MEM_BUDGET_MB", "12")
    name, seq, seeds = case.reads[0]
    bad = [("noseeds", seq, []), ("badnode", seq, [(999999, 0, False)]), ("ok", seq, seeds)]
    pipe = api.Pipeline(api.Graph.from_case(case), depth=2)
    out = [r.as_dicts() for r in pipe.align_all([case.reads, [], bad, case.reads, []], case.b, case.B)]
    assert [len(x) for x in out] == [len(case.reads), 0, 3, len(case.reads), 0]
    assert_same(out[0], expected, "pipeline, split batch 0")
    assert_same(out[3], expected, "pipeline, split batch 3")
    assert [x["failed"] for x in out[2]] == [1, 1, 0]
    assert out[2][2]["score"] == expected[0]["score"]
    assert pipe.stats()["launches"] >= 8
    pipe.close()


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed,rate", [(201, 0.05), (202, 0.4)])
def test_iupac_and_lower_case_reads(api, tmp_path, seed, rate):
    # ambiguity codes and lower-case letters in the reads: IUPAC matching inside a slice (GraphAligner.h:2039-2110), exact
    # compare on the row above a slice (:1503,1540), ReverseComplement of the backward part (CommonUtils.cpp:60-136)
    import numpy as np
    rng = np.random.default_rng(seed)
    g = synth.make_graph(seed, 8000, chop=32, bubble_every=80, inversion_every=700)
    case = synth.make_case(seed, g, 16, 700, b=10, seed_offsets=(0, 350), errors=(0.03, 0.03, 0.03))
    codes = "RYSWKMBDHVNryswkmbdhvn"
    reads = []
    for name, seq, seeds in case.reads:
        s = list(seq)
        for i in range(len(s)):
            if rng.random() < rate:
                s[i] = codes[int(rng.integers(0, len(codes)))] if rng.random() < 0.7 else s[i].lower()
        reads.append((name, "".join(s), seeds))
    case.reads = reads
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=4)
    aligner = api.Aligner(api.Graph.from_case(case))
    assert_same(aligner.align(case.reads, case.b, case.B).as_dicts(), expected, "iupac %g" % rate)
    aligner.close()


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
def test_reads_shorter_than_one_slice(api, tmp_path):
    # 1 .. 70 bp: a single 64-row slice of mostly padding; seeds on the first, a middle and the last base
    import numpy as np
    rng = np.random.default_rng(203)
    g = synth.make_graph(203, 3000, chop=16, bubble_every=60)
    reads = []
    for k, ln in enumerate([1, 2, 3, 7, 31, 32, 33, 63, 64, 65, 70]):
        read, real, walk, mp = synth.simulate_read(rng, g, max(2, ln), 0.03, 0.03, 0.03)
        read = read[:ln]
        for off in sorted(set([0, len(read) // 2, len(read) - 1])):
            reads.append(("r%d_%d" % (k, off), read, synth.seeds_for(walk, mp, len(read), [min(off, len(mp) - 1)])))
    case = gacase.Case(list(g.nodes), list(g.edges), reads, 5, 0)
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=4)
    aligner = api.Aligner(api.Graph.from_case(case))
    assert_same(aligner.align(case.reads, case.b, case.B).as_dicts(), expected, "tiny reads")
    aligner.close()


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
@pytest.mark.parametrize("overlap", [1, 3, 5])
def test_gfa_edge_overlap(api, tmp_path, overlap):
    # GFA with a k-bp edge overlap: every node loses its last k bases (BigraphToDigraph.cpp:58-67) and the backward part of
    # a split read is extended by DBGOverlap (GraphAligner.h:2991-2992); mid-read seeds exercise the latter
    g = synth.make_graph(210 + overlap, 10000, chop=16)
    case = synth.make_case(210 + overlap, g, 12, 600, b=10, seed_offsets=(0, 200, -60), errors=(0.02, 0.02, 0.02))
    case.gfa_overlap = overlap
    path = str(tmp_path / "case.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=4)
    aligner = api.Aligner(api.Graph.from_case(case))
    assert_same(aligner.align(case.reads, case.b, case.B).as_dicts(), expected, "gfa overlap %d" % overlap)
    aligner.close()


@pytest.mark.parametrize("name", ["bubbles_multiseed", "seed_pos1", "ragged_short"])
def test_page_locked_input_is_used_in_place(api, golden_dir, name, monkeypatch):
    # reads handed over in page-locked host memory are uploaded from the caller's buffer (no staging copy); the results must
    # not depend on where the bytes lie: pinned, pinned through the pipeline, and the staged path forced by the environment
    import torch
    case = gacase.read_case(os.path.join(golden_dir, name + ".gacase"))
    expected = load_expected(os.path.join(golden_dir, name + ".expected"))
    held = []

    def pinned(nbytes):
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        held.append(t)
        return t.numpy()

    names = [r[0] for r in case.reads]
    graph = api.Graph.from_case(case)
    aligner = api.Aligner(graph)
    packed = api.PackedReads(case.reads, case.b, case.B, seq_alloc=pinned)
    res = aligner.align(packed)
    res.names = names
    assert_same(res.as_dicts(), expected, name + " (pinned input)")
    res.free()
    monkeypatch.setenv("GA_NO_INPLACE_INPUT", "1")
    res = aligner.align(packed)
    res.names = names
    assert_same(res.as_dicts(), expected, name + " (pinned input, staged)")
    res.free()
    monkeypatch.delenv("GA_NO_INPLACE_INPUT")
    aligner.close()
    pipe = api.Pipeline(graph, depth=2)
    for r in pipe.align_all([packed] * 3):
        r.names = names
        assert_same(r.as_dicts(), expected, name + " (pinned input, pipeline)")
        r.free()
    pipe.close()


@pytest.mark.parametrize("budget", ["1", "3"])
def test_seed_rounds_with_split_launches(api, golden_dir, budget, monkeypatch):
    # reads with several seeds run in two rounds (first seed, then the seeds its alignment does not cover); with a tiny device
    # budget each round is cut into several launches and the rounds' results are merged read by read - blocking call and pipeline
    monkeypatch.setenv("GA_MEM_BUDGET_MB", budget)
    for name in ("bubbles_multiseed", "seed_pos1"):
        case = gacase.read_case(os.path.join(golden_dir, name + ".gacase"))
        expected = load_expected(os.path.join(golden_dir, name + ".expected"))
        graph = api.Graph.from_case(case)
        aligner = api.Aligner(graph)
        res = aligner.align(case.reads, case.b, case.B)
        assert_same(res.as_dicts(), expected, name)
        assert aligner.stats()["launches"] >= 4   # more than one launch sequence (two rounds, each possibly split)
        res.free()
        aligner.close()
        pipe = api.Pipeline(graph, depth=2)
        for r in pipe.align_all([api.PackedReads(case.reads, case.b, case.B)] * 2):
            r.names = [x[0] for x in case.reads]
            assert_same(r.as_dicts(), expected, name + " (pipeline)")
            r.free()
        pipe.close()


def test_all_seeds_at_once_gives_the_same_results(api, golden_dir, tmp_path):
    # the two-round seed loop against the one-round form (every seed aligned speculatively, the pruning replayed afterwards), in a
    # fresh process because the switch is read once
    import subprocess
    import sys
    script = ("import os, sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
              "from graphaligner_b200 import api\nfrom graphaligner_b200.tools import gacase\nfrom helpers import assert_same, load_expected\n"
              "case = gacase.read_case(%r); exp = load_expected(%r)\n"
              "al = api.Aligner(api.Graph.from_case(case)); res = al.align(case.reads, case.b, case.B)\n"
              "assert_same(res.as_dicts(), exp, 'one round'); print('streams', al.stats()['streams'])\n"
              % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
                 os.path.join(golden_dir, "bubbles_multiseed.gacase"), os.path.join(golden_dir, "bubbles_multiseed.expected")))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=dict(os.environ, GA_ALL_SEEDS_AT_ONCE="1"))
    assert r.returncode == 0, r.stderr[-800:]
