import os
import subprocess

from graphaligner_b200.tools import gacase

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ALIGN = os.path.join(ROOT, "oracle", "_ref", "ref_align")
GOLDEN = ["smallexample", "dag_snp", "bubbles_multiseed", "tangle_cycles", "seed_pos1", "gfa", "ragged_short", "wide_band", "ramp", "ramp_redo", "ramp_stale", "ramp_stale_long", "cyclic_partial_confirm", "cyclic_last_call_min"]
# fixtures whose expected output depends on the reference's work-list schedule inside cyclic components; the cell-by-cell
# restatement (oracle/ga_oracle.cpp) evaluates the fix point and says so in its header
SCHEDULE_DEPENDENT = ["cyclic_partial_confirm", "cyclic_last_call_min"]
# fixtures whose expected output comes from a stale sqrt checkpoint after a -B ramp redo (GraphAligner.h:2667,2772-2786,2858-2943)
STALE_CHECKPOINT = ["ramp_stale", "ramp_stale_long"]
KEYS = ("failed", "score", "start", "end", "qpos", "nmap", "ntrace", "th")


def load_expected(path):
    with open(path) as f:
        reads, _ = gacase.parse_ref_output(f.read())
    return reads


def run_reference(case_path, threads=1, extra=()):
    """Runs the UNMODIFIED reference hot path (oracle/_ref/ref_align, test infrastructure only)."""
    res = subprocess.run([REF_ALIGN, case_path, "--quiet", "--threads", str(threads)] + list(extra), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("reference oracle failed (rc %d): %s" % (res.returncode, res.stderr[-500:]))
    return gacase.parse_ref_output(res.stdout)


def assert_same(mine, expected, what=""):
    assert len(mine) == len(expected), what
    for m, e in zip(mine, expected):
        for k in KEYS:
            assert m[k] == e[k], "%s read %s: %s differs: %r != %r" % (what, e["name"], k, m[k], e[k])
        assert [tuple(x) for x in m["mappings"]] == [tuple(x) for x in e["mappings"]], "%s read %s: mappings differ" % (what, e["name"])
