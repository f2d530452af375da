"""Config-2 sized parity: properties that do not depend on the input size, plus a sample checked against the reference.

BASELINE config 2 = 5 Mbp graph (32-bp nodes, SNP bubble / 1000 bp), 10 kbp reads at ~15 % error, band 10, seed at offset 0.
The graph is full size; 2 000 reads keep the test at about a minute (bench.py runs all 10 000)."""
import os

import numpy as np
import pytest

from graphaligner_b200.tools import gacase, synth
from helpers import REF_ALIGN, assert_same, run_reference

pytestmark = pytest.mark.gpu

N_READS = 2000


@pytest.fixture(scope="module")
def workload():
    from graphaligner_b200 import api
    g = synth.make_graph(1, 5_000_000, chop=32, snp_every=1000)
    case = synth.make_case(1000, g, N_READS, 10_000, b=10)
    graph = api.Graph.from_case(case)
    aligner = api.Aligner(graph)
    packed = api.PackedReads(case.reads, 10, 0)
    res = aligner.align(packed)
    yield api, case, aligner, packed, res
    aligner.close()


def test_every_read_aligns_and_invariants_hold(workload):
    api, case, aligner, packed, res = workload
    r = res.reads
    assert int((r["flags"] & 1).sum()) == 0, "a DP stream hit a hard limit"
    ok = r["failed"] == 0
    assert ok.mean() > 0.99   # WHICH reads fail is pinned against the reference in test_every_read_against_reference
    lens = np.array([len(x[1]) for x in case.reads])
    # alignmentEnd - alignmentStart = 64 * (retained slices), never more than the padded read (GraphAligner.h:486)
    span = (r["alignment_end"] - r["alignment_start"])[ok]
    assert np.all(span % 64 == 0) and np.all(span <= ((lens[ok] + 63) // 64) * 64)
    # unit-cost edit distance of a ~15 % error read: positive, far below the read length
    assert np.all(r["score"][ok] > 0) and np.all(r["score"][ok] < 0.3 * lens[ok])
    # seed at offset 0: no backward part, query position 0
    assert np.all(r["query_position"][ok] == 0)
    # forward word updates: one per band column and slice, never fewer than one column per slice
    assert np.all(r["word_columns"][ok] >= span // 64)


def test_paths_are_walks_in_the_graph(workload):
    # consecutive mappings must be joined by an edge of the digraph (checked on a sample through the case's edge list)
    api, case, aligner, packed, res = workload
    succ = synth.Graph()
    succ.nodes, succ.edges = case.nodes, case.edges
    s = succ.succ()
    d = res.as_dicts()
    for x in d[:200]:
        if x["failed"]:
            continue
        nodes = [(m[0] // 2, m[0] % 2) for m in x["mappings"]]
        for a, b in zip(nodes, nodes[1:]):
            assert b in s.get(a, []), "mapping %s -> %s is not an edge" % (a, b)
        # to_lengths cover the aligned part of the read exactly once
        assert sum(m[4] for m in x["mappings"]) <= len(case.reads[d.index(x)][1])


def test_results_do_not_depend_on_batch_composition(workload, monkeypatch):
    # the same reads aligned alone, in a different order, and in a batch split into several launches give identical results
    api, case, aligner, packed, res = workload
    full = res.as_dicts()
    idx = list(range(0, N_READS, 97))
    sub = [case.reads[i] for i in reversed(idx)]
    a = aligner.align(sub, 10, 0).as_dicts()
    for k, i in enumerate(reversed(idx)):
        for key in ("failed", "score", "start", "end", "nmap", "ntrace", "th"):
            assert a[k][key] == full[i][key], (i, key)
    monkeypatch.setenv("GA_MEM_BUDGET_MB", "600")
    chunked = aligner.align(case.reads[:400], 10, 0).as_dicts()
    for i in range(400):
        for key in ("failed", "score", "start", "end", "nmap", "ntrace", "th"):
            assert chunked[i][key] == full[i][key], (i, key)


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
def test_every_read_against_reference(workload, tmp_path):
    # ALL reads of the full-size batch through the reference (ref_align --summary: one line per read with the trace fingerprint
    # and an order-sensitive checksum of the mappings): the set of failed reads must be the reference's set (a read fails when
    # the correctness HMM rejects every slice, GraphAligner.h:2554-2569, e.g. a seed on the wrong strand of a repeat), and every
    # other read must agree in score, range, query position, mapping count + checksum and trace-item count + fingerprint
    api, case, aligner, packed, res = workload
    path = str(tmp_path / "full.gacase")
    gacase.write_case(case, path)
    expected, _ = run_reference(path, threads=os.cpu_count() or 8, extra=["--summary"])
    assert len(expected) == N_READS
    r = res.reads
    ref_failed = sorted(i for i, e in enumerate(expected) if e["failed"])
    my_failed = sorted(int(i) for i in np.nonzero(r["failed"])[0])
    assert my_failed == ref_failed, "failed reads differ: here %s, reference %s" % (my_failed[:10], ref_failed[:10])
    mh = gacase.mapping_checksums(r, res.mappings)
    ok = [i for i in range(N_READS) if not expected[i]["failed"]]
    for key, col in (("score", "score"), ("start", "alignment_start"), ("end", "alignment_end"), ("qpos", "query_position"), ("nmap", "n_mappings"), ("ntrace", "n_trace")):
        mine = r[col][ok].astype(np.int64)
        ref = np.array([expected[i][key] for i in ok], dtype=np.int64)
        bad = np.nonzero(mine != ref)[0]
        assert len(bad) == 0, "%s differs for %d reads, first: read %d here %d reference %d" % (key, len(bad), ok[bad[0]], mine[bad[0]], ref[bad[0]])
    ref_mh = np.array([expected[i]["mh"] for i in ok], dtype=np.uint64)
    assert np.array_equal(mh[ok], ref_mh), "mapping checksums differ for %d reads" % int((mh[ok] != ref_mh).sum())
    # trace fingerprints (materialised per read on the host): every 7th read plus the first and last hundred
    for i in sorted(set(ok[::7]) | set(ok[:100]) | set(ok[-100:])):
        assert res.trace_hash(i) == expected[i]["th"], "trace fingerprint of read %d" % i
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "full_size_failed_reads.txt") if os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")) else os.devnull, "w") as f:
        f.write("config 2, %d reads: failed here %s; failed in the reference %s\n" % (N_READS, my_failed, ref_failed))
