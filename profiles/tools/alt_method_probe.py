"""One more probe of the reference's alternate method: -B > -b and a band that
crosses GraphAlignerCommon.h:10's 200 000-bp cutoff only in a MIDDLE slice.  Three nodes A (50 kbp) -> B (101 kbp) ->
C (101 kbp); the read is the tail of A, all of B and the head of C, seeded on A: the band is A + B (151 kbp, bit-vector
method) until the read nears the end of B, then B + C (202 kbp, calculateSliceAlternate).  Prints what the two flavours of
the unmodified reference (oracle/_ref/ref_align = -DNDEBUG, ref_align_stock = asserts on) do with it.
    python profiles/tools/alt_method_probe.py"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphaligner_b200.tools import gacase  # noqa: E402

rng = np.random.default_rng(7)
seq = lambda n: "".join("ACGT"[i] for i in rng.integers(0, 4, n))
a, b, c = seq(50_000), seq(101_000), seq(101_000)
read = a[-200:] + b + c[:300]
for bw, ramp in ((10, 20), (10, 0), (30, 60)):
    case = gacase.Case([(2, a), (3, b), (4, c)], [(2, False, 3, False), (3, False, 4, False)], [("crossmid", read, [(2, 0, False)])], bw, ramp)
    path = "/tmp/alt_probe.gacase"
    gacase.write_case(case, path)
    for exe in ("ref_align", "ref_align_stock"):
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe), path, "--quiet", "--threads", "1"], capture_output=True, text=True)
        reads = [l for l in r.stdout.split("\n") if l.startswith("READ")]
        print("b=%d B=%d %-16s rc=%d %s %s" % (bw, ramp, exe, r.returncode, reads[0][:120] if reads else "(no result)", r.stderr.strip().split("\n")[-1][:160] if r.returncode else ""), flush=True)
