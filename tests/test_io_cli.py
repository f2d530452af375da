"""Host surface around the hot path: vg / GFA loaders, the GAM codec and the reference-compatible command line."""
import os
import subprocess

import pytest

from graphaligner_b200.tools import gacase, synth, vgio
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


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,batch_bp,devices", [("smallexample", None, None), ("bubbles_multiseed", None, None), ("gfa", None, None), ("bubbles_multiseed", 5000, None),
                                                   ("bubbles_multiseed", 5000, "0,0"), ("bubbles_multiseed", None, "0-1"), ("dag_snp", 3000, "0,1"),
                                                   ("bubbles_multiseed", "per-read", None)])
def test_cli_end_to_end(tmp_path, golden_dir, name, batch_bp, devices):
    # batch_bp: the driver streams the read set through two contexts per device in batches of about that many read bases
    # devices: -G list (the counterpart of the reference's worker threads, Aligner.cpp:285-306); "0,0" runs the batch queue
    # with four lanes on one GPU, the others need two GPUs
    # batch_bp "per-read": -t 2 worker threads calling the per-read AlignOneWay concurrently, as the reference's driver does
    # (Aligner.cpp:107-140)
    if devices and devices != "0,0" and _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    extra_env = {}
    if batch_bp == "per-read":
        extra_env["GA_PER_READ"] = "1"
    elif batch_bp:
        extra_env["GA_BATCH_BP"] = str(batch_bp)
    case = gacase.read_case(os.path.join(golden_dir, name + ".gacase"))
    expected = {e["name"]: e for e in load_expected(os.path.join(golden_dir, name + ".expected"))}
    graph_path, fastq, seeds = vgio.write_case_files(case, str(tmp_path / "c"))
    out = str(tmp_path / "out.gam")
    aug = str(tmp_path / "aug.vg")
    r = subprocess.run([ALIGNER, "-g", graph_path, "-f", fastq, "-s", seeds, "-a", out, "-t", "2", "-b", str(case.b)] + (["-B", str(case.B)] if case.B else [])
                       + (["-G", devices] if devices else []) + (["-A", aug] if graph_path.endswith(".vg") else []),
                       capture_output=True, text=True, cwd=str(tmp_path), env=dict(os.environ, **extra_env))
    assert r.returncode == 0, r.stderr[-500:]
    if graph_path.endswith(".vg"):
        # -A (augmentGraphwithAlignment, Aligner.cpp:24-74): the input graph's nodes, one edge per pair of consecutive mappings
        in_nodes, _ = vgio.load_vg_graph(graph_path)
        aug_nodes, aug_edges = vgio.load_vg_graph(aug)
        assert [(n["id"], n["sequence"], n["name"]) for n in aug_nodes] == [(n["id"], n["sequence"], n["name"]) for n in in_nodes]
        want_edges = []
        for a in vgio.load_gam(out):
            p = a["path"]
            want_edges += [(p[i]["position"]["node_id"], p[i + 1]["position"]["node_id"], bool(p[i]["position"]["is_reverse"]), bool(p[i + 1]["position"]["is_reverse"]))
                           for i in range(len(p) - 1)]
        assert [(e["from"], e["to"], e["from_start"], e["to_end"]) for e in aug_edges] == want_edges
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


def test_evaluation_tools_roundtrip(tmp_path):
    # SimulateReads / PickSeedHits / CompareAlignments equivalents over the GAM codec, no GPU involved:
    # the truth compared with itself is all good matches; dropping a read or swapping its path makes it a bad match
    from graphaligner_b200.tools import evaltools
    g = synth.make_graph(21, 20000, chop=32, bubble_every=200)
    graph_path = str(tmp_path / "g.vg")
    vgio.write_stream(graph_path, [vgio.encode_graph(g.nodes, g.edges)])
    truth, fastq, seeds = str(tmp_path / "truth.gam"), str(tmp_path / "reads.fastq"), str(tmp_path / "seeds.gam")
    assert evaltools.main(["evaltools", "simulate", graph_path, truth, fastq, "12", "1500", "0.05", "0.05", seeds, "0.05"]) == 0
    reads = vgio.load_fastq(fastq)
    t = vgio.load_gam(truth)
    assert len(reads) == 12 and [a["name"] for a in t] == [n for n, _ in reads]
    assert all(len(a["sequence"]) == 1500 and len(a["path"]) >= 1500 // 33 for a in t)
    # pickseeds: duplicates and surplus hits are dropped, ids <= 1 ignored, output grouped by sorted name
    extra = str(tmp_path / "extra.gam")
    first = t[0]["name"]
    vgio.write_stream(extra, [vgio.encode_seed(first, 1, 5, False), vgio.encode_seed(first, 77, 9, False), vgio.encode_seed(first, 78, 9, True),
                              vgio.encode_seed(first, 77, 9, True)])
    picked = str(tmp_path / "picked.gam")
    assert evaltools.main(["evaltools", "pickseeds", picked, "2", seeds, seeds, extra]) == 0
    p = vgio.load_gam(picked)
    assert [a["name"] for a in p] == sorted([n for n, _ in reads] + [first])
    mine = [a for a in p if a["name"] == first]
    assert len(mine) == 2 and mine[1]["path"][0]["position"]["node_id"] == 77 and mine[1]["query_position"] == 9
    import io
    out = io.StringIO()
    assert evaltools.compare_alignments(truth, truth, graph_path, out) == (12, 0)
    assert "good matches: 12" in out.getvalue()
    # a prediction on other nodes is a bad match; a missing one too
    other = g.nodes[-1][0]
    bad = [evaltools.encode_alignment(t[0]["name"], t[0]["sequence"], [(other, 0, False)], score=3)]
    bad += [evaltools.encode_alignment(a["name"], a["sequence"], [(m["position"]["node_id"], 0, m["position"]["is_reverse"]) for m in a["path"]]) for a in t[2:]]
    pred = str(tmp_path / "pred.gam")
    vgio.write_stream(pred, bad)
    assert evaltools.compare_alignments(truth, pred, graph_path, io.StringIO()) == (10, 2)


@pytest.mark.gpu
def test_simulate_align_compare_end_to_end(tmp_path):
    # the reference's evaluation loop: SimulateReads -> Aligner -> CompareAlignments; every simulated read must land on its
    # true path (identity >= 0.7, CompareAlignments.cpp:83)
    from graphaligner_b200.tools import evaltools
    g = synth.make_graph(22, 60000, chop=32, bubble_every=300)
    graph_path = str(tmp_path / "g.vg")
    vgio.write_stream(graph_path, [vgio.encode_graph(g.nodes, g.edges)])
    truth, fastq, seeds, out = (str(tmp_path / n) for n in ("truth.gam", "reads.fastq", "seeds.gam", "out.gam"))
    assert evaltools.main(["evaltools", "simulate", graph_path, truth, fastq, "40", "3000", "0.05", "0.05", seeds, "0.05"]) == 0
    r = subprocess.run([ALIGNER, "-g", graph_path, "-f", fastq, "-s", seeds, "-a", out, "-t", "2", "-b", "10"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-500:]
    good, bad = evaltools.compare_alignments(truth, out, graph_path, open(os.devnull, "w"))
    assert good >= 38 and good + bad == 40
