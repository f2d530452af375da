"""Reads with IUPAC ambiguity codes and lower-case letters (characterMatch GraphAligner.h:2039-2110, ReverseComplement
CommonUtils.cpp:60-136, and the exact-compare previousEq quirk GraphAligner.h:1503,1540) against the reference run on the box.
python profiles/tools/iupac_probe.py FIRST COUNT"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graphaligner_b200 import api
from graphaligner_b200.tools import fuzz, gacase, synth

CODES = "RYSWKMBDHVNryswkmbdhvnacgt"
api.load_library()
first, count = int(sys.argv[1]), int(sys.argv[2])
same = differ = crashed = 0
for it in range(first, first + count):
    rng = np.random.default_rng(it)
    g = synth.make_graph(it, 8000, chop=int(rng.choice([8, 32, 64])), bubble_every=int(rng.integers(20, 200)), inversion_every=int(rng.choice([0, 700])))
    rl = int(rng.choice([100, 400, 1500]))
    case = synth.make_case(it, g, 12, rl, b=int(rng.choice([5, 10, 30])), seed_offsets=[(0,), (rl // 2,), (0, -40)][int(rng.integers(0, 3))], errors=(0.03, 0.03, 0.03))
    rate = float(rng.choice([0.02, 0.1, 0.4]))
    reads = []
    for name, seq, seeds in case.reads:
        s = list(seq)
        for i in range(len(s)):
            if rng.random() < rate:
                c = CODES[int(rng.integers(0, len(CODES)))]
                s[i] = c if c not in "acgt" else s[i].lower()
        reads.append((name, "".join(s), seeds))
    case.reads = reads
    path = "/tmp/iupac_%d.gacase" % it
    gacase.write_case(case, path)
    ref = subprocess.run([fuzz.REF, path, "--quiet", "--threads", "2"], capture_output=True, text=True)
    if ref.returncode != 0:
        crashed += 1
        continue
    expected, _ = gacase.parse_ref_output(ref.stdout)
    al = api.Aligner(api.Graph.from_case(case))
    mine = al.align(case.reads, case.b, 0).as_dicts()
    al.close()
    bad = [e["name"] for m, e in zip(mine, expected) if any(m[k] != e[k] for k in fuzz.KEYS) or [tuple(x) for x in m["mappings"]] != [tuple(x) for x in e["mappings"]]]
    if bad:
        differ += 1
        print("DIFF", it, bad[:4], "rate", rate, "rl", rl, flush=True)
    else:
        same += 1
print("iupac reads %d..%d: identical %d, different %d, reference crashed %d" % (first, first + count - 1, same, differ, crashed))
