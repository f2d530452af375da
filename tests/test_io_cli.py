"""Host surface around the hot path: vg / GFA loaders, the GAM codec and the reference-compatible command line."""
import os
import subprocess

import pytest

from graphaligner_b200.tools import gacase, vgio
from helpers import ROOT, load_expected

ALIGNER = os.path.join(ROOT, "graphaligner_b200", "bin", "Aligner")


def test_vg_loader_matches_in_memory_builder(lib_built, tmp_path, golden_dir):
    from graphaligner_b200 import api
    case = gacase.read_case(os.path.join(golden_dir, "bubbles_multiseed.gacase"))
    graph_path, _, _ = vgio.write_case_files(case, str(tmp_path / "c"))
    a = api.Graph.from_case(case)
    b = api.Graph.load(graph_path)
    assert (a.node_count(), a.size_bp(), a.edge_count()) == (b.node_count(), b.size_bp(), b.edge_count())
    # and the python decoder reads back what the python encoder wrote
    n, e = vgio.load_vg_graph(graph_path)
    assert [(x["id"], x["sequence"]) for x in n] == case.nodes


def test_gfa_loader_matches_in_memory_builder(lib_built, tmp_path, golden_dir):
    from graphaligner_b200 import api
    case = gacase.read_case(os.path.join(golden_dir, "gfa.gacase"))
    graph_path, _, _ = vgio.write_case_files(case, str(tmp_path / "c"))
    a = api.Graph.from_case(case)
    b = api.Graph.load(graph_path)
    assert (a.node_count(), a.size_bp(), a.edge_count()) == (b.node_count(), b.size_bp(), b.edge_count())


def test_reference_smallexample_files_load(lib_built):
    # the reference's own shipped graph, when present (build container only)
    path = "/root/reference/test/smallexample/sub_test.vg"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    from graphaligner_b200 import api
    g = api.Graph.load(path)
    assert (g.node_count(), g.size_bp(), g.edge_count()) == (38, 332, 50)


def test_cli_validation_messages():
    if not os.path.exists(ALIGNER):
        pytest.skip("Aligner not built")
    # AlignerMain.cpp:68-96: message on stderr, exit code 0
    r = subprocess.run([ALIGNER, "-g", "x.vg", "-f", "r.fq", "-t", "1", "-b", "1", "-s", "s.gam"], capture_output=True, text=True)
    assert r.returncode == 0 and "bandwidth must be >= 2" in r.stderr
    r = subprocess.run([ALIGNER, "-g", "x.vg", "-f", "r.fq", "-t", "1", "-b", "10"], capture_output=True, text=True)
    assert "either initial full band or seed file must be set" in r.stderr
    r = subprocess.run([ALIGNER, "-g", "x.vg", "-f", "r.fq", "-t", "0", "-b", "10", "-s", "s.gam"], capture_output=True, text=True)
    assert "number of threads must be >= 1" in r.stderr
    r = subprocess.run([ALIGNER, "-g", "x.vg", "-f", "r.fq", "-t", "1", "-b", "10", "-B", "5", "-s", "s.gam"], capture_output=True, text=True)
    assert "backup bandwidth must be higher than initial bandwidth" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,batch_bp", [("smallexample", None), ("bubbles_multiseed", None), ("gfa", None), ("bubbles_multiseed", 5000)])
def test_cli_end_to_end(tmp_path, golden_dir, name, batch_bp):
    # batch_bp: the driver streams the read set through two contexts in batches of about that many read bases
    case = gacase.read_case(os.path.join(golden_dir, name + ".gacase"))
    expected = {e["name"]: e for e in load_expected(os.path.join(golden_dir, name + ".expected"))}
    graph_path, fastq, seeds = vgio.write_case_files(case, str(tmp_path / "c"))
    out = str(tmp_path / "out.gam")
    r = subprocess.run([ALIGNER, "-g", graph_path, "-f", fastq, "-s", seeds, "-a", out, "-t", "2", "-b", str(case.b)] + (["-B", str(case.B)] if case.B else []),
                       capture_output=True, text=True, cwd=str(tmp_path), env=dict(os.environ, **({"GA_BATCH_BP": str(batch_bp)} if batch_bp else {})))
    assert r.returncode == 0, r.stderr[-500:]
    alns = vgio.load_gam(out)
    ok = [e for e in expected.values() if not e["failed"]]
    assert len(alns) == len(ok)
    assert "final result has %d alignments" % len(ok) in r.stderr
    # reads are consumed last to first (Aligner.cpp:111-117)
    assert [a["name"] for a in alns] == [n for n, _, _ in reversed(case.reads) if not expected[n]["failed"]]
    for a in alns:
        e = expected[a["name"]]
        assert a["score"] == e["score"] and a["query_position"] == e["qpos"]
        # the driver halves the digraph node ids (Aligner.cpp:83-91)
        got = [(m["position"]["node_id"], int(m["position"]["is_reverse"]), m["position"]["offset"], m["edits"][0]["from_length"], m["edits"][0]["to_length"]) for m in a["path"]]
        want = [(nid // 2, rev, off, fl, tl) for nid, rev, off, fl, tl in e["mappings"]]
        assert got == want
        # each edit carries its slice of the read
        seq = dict((n, s) for n, s, _ in case.reads)[a["name"]]
        assert a["sequence"] == seq
    first = ok[0]["name"]
    assert os.path.exists(str(tmp_path / ("alignment_0_%s.gam" % first))) and os.path.exists(str(tmp_path / ("trace_0_%s.trace" % first)))
    assert "read %s score %d" % (first, ok[0]["score"]) in r.stdout
