"""CPU-side checks (no GPU): the C-ABI library loads and exports every declared symbol, the host graph
builder matches the reference's graph statistics, and the oracle reproduces the committed golden vectors."""
import os
import re

import pytest

from graphaligner_b200.tools import gacase
from helpers import GOLDEN, REF_ALIGN, SCHEDULE_DEPENDENT, STALE_CHECKPOINT, ROOT, assert_same, load_expected, run_reference


def test_library_exports_every_declared_symbol(lib_built):
    from graphaligner_b200 import api
    header = open(os.path.join(ROOT, "include", "graphaligner_b200.h")).read()
    declared = set(re.findall(r"\b(ga_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib_built, name), "missing export " + name
    assert declared == set(api.EXPORTS)


def test_no_gpu_fails_loudly(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from graphaligner_b200 import api
    case = gacase.read_case(os.path.join(ROOT, "tests", "golden", "smallexample.gacase"))
    g = api.Graph.from_case(case)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        api.Aligner(g)


def test_graph_builder_counts(lib_built):
    # reference Finalize prints 38 nodes / 332bp / 50 edges for test/smallexample (oracle/_ref/ref_align_stock)
    from graphaligner_b200 import api
    case = gacase.read_case(os.path.join(ROOT, "tests", "golden", "smallexample.gacase"))
    g = api.Graph.from_case(case)
    assert (g.node_count(), g.size_bp(), g.edge_count()) == (38, 332, 50)


@pytest.mark.skipif(not os.path.exists(REF_ALIGN), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", GOLDEN)
def test_reference_oracle_reproduces_golden(golden_dir, name):
    expected = load_expected(os.path.join(golden_dir, name + ".expected"))
    got, _ = run_reference(os.path.join(golden_dir, name + ".gacase"))
    assert_same(got, expected, name)


RESTATEMENT = os.path.join(ROOT, "oracle", "_ref", "ga_oracle")


@pytest.mark.skipif(not os.path.exists(RESTATEMENT), reason="oracle restatement not built (python __graft_entry__.py)")
@pytest.mark.parametrize("name", [n for n in GOLDEN if n not in SCHEDULE_DEPENDENT])
def test_restatement_oracle_reproduces_golden(golden_dir, name):
    # oracle/ga_oracle.cpp (cell-by-cell restatement) pinned against outputs of the unmodified reference
    import subprocess
    res = subprocess.run([RESTATEMENT, os.path.join(golden_dir, name + ".gacase")], capture_output=True, text=True, check=True)
    got, _ = gacase.parse_ref_output(res.stdout)
    assert_same(got, load_expected(os.path.join(golden_dir, name + ".expected")), name)


@pytest.mark.skipif(not (os.path.exists(RESTATEMENT) and os.path.exists(REF_ALIGN)), reason="oracle not built")
def test_restatement_oracle_ramp_redo_against_reference_run_here(tmp_path):
    # -B ramp redo with the reference's sqrt checkpoints (GraphAligner.h:2648-2719,2772-2786,2858-2943) in the cell-by-cell restatement:
    # ramp fuzz cases on acyclic graph kinds, every case the reference survives must be identical (the restatement stops with
    # exit code 8 where the reference reads a slice that is not there)
    import subprocess
    from graphaligner_b200.tools import fuzz
    assert set(STALE_CHECKPOINT) <= set(GOLDEN)
    fuzz.RAMP = True
    try:
        cases = [(seed, fuzz.make_case(seed)) for seed in range(30000, 30060) if seed % 6 in (0, 1, 2)]
    finally:
        fuzz.RAMP = False
    compared = 0
    for seed, (case, desc) in cases:
        if desc["rl"] > 1000:
            continue
        path = str(tmp_path / ("r%d.gacase" % seed))
        gacase.write_case(case, path)
        try:
            expected, _ = run_reference(path)
        except RuntimeError:
            continue
        res = subprocess.run([RESTATEMENT, path], capture_output=True, text=True)
        assert res.returncode == 0, "seed %d: %s" % (seed, res.stderr[-300:])
        got, _ = gacase.parse_ref_output(res.stdout)
        assert_same(got, expected, "restatement, ramp fuzz %d" % seed)
        compared += 1
    assert compared >= 8


def test_parallel_generators_make_valid_cases(tmp_path):
    # bench.py's full-size configs build graph and reads on worker processes (segments joined exit -> entry, chunked read
    # seeds): the reference must align reads that cross segment borders, and the read set must not depend on the worker count
    import os
    import pytest
    from graphaligner_b200.tools import gacase, synth
    from helpers import REF_ALIGN, run_reference
    g = synth.make_graph_parallel(11, 6000, workers=2, segment=1500, chop=32, bubble_every=100, indel_frac=0.2)
    assert len(g.exits) >= 1 and len(g.nodes) == len(g.seq)
    a = synth.make_case_parallel(5, g, 24, 3000, workers=2, chunk=8, b=10)
    b = synth.make_case_parallel(5, g, 24, 3000, workers=1, chunk=8, b=10)
    assert a.reads == b.reads and [r[0] for r in a.reads] == ["read_%d" % i for i in range(24)]
    shard = synth.make_case_parallel(5, g, 24, 3000, workers=2, chunk=8, read_range=(8, 16), b=10)
    assert shard.reads == a.reads[8:16]
    if not os.path.exists(REF_ALIGN):
        pytest.skip("oracle/_ref not built")
    path = str(tmp_path / "par.gacase")
    gacase.write_case(a, path)
    expected, _ = run_reference(path, threads=4)
    assert sum(1 for e in expected if not e["failed"]) >= 22
    assert all(e["score"] < 0.3 * 3000 for e in expected if not e["failed"])


def test_word_level_primitives_against_reference_wordslice():
    # the column primitives of the product (ga_merge_cols, ga_col_value, ga_vertical_merge: ga_core.cuh compiled
    # for the host) against the reference's own WordSlice::mergeWith / getValue (WordSlice.h:202,223,361-421) on random columns
    import os
    import subprocess
    import pytest
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "word_harness")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/word_harness not built")
    for seed in (1, 2):
        r = subprocess.run([exe, "100000", str(seed)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-600:]
        assert "100000 merges" in r.stdout and "identical" in r.stdout
