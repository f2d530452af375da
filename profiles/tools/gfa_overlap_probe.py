"""GFA graphs with an edge overlap of 1 .. 5 bp (ConvertGFANodeToNodes trims every node, BigraphToDigraph.cpp:58-67; the backward
part of a split read is extended by DBGOverlap, GraphAligner.h:2991-2992) against the reference run on the box.
python profiles/tools/gfa_overlap_probe.py FIRST COUNT"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from graphaligner_b200 import api
from graphaligner_b200.tools import fuzz, gacase, synth

api.load_library()
first, count = int(sys.argv[1]), int(sys.argv[2])
same = differ = crashed = 0
for it in range(first, first + count):
    rng = np.random.default_rng(it)
    g = synth.make_graph(it, 10000, chop=int(rng.choice([8, 16, 32, 64])), snp_every=0, bubble_every=0)
    rl = int(rng.choice([150, 600, 2000]))
    case = synth.make_case(it, g, 10, rl, b=int(rng.choice([5, 10, 30])), seed_offsets=[(0,), (rl // 2,), (0, rl // 3, -60)][int(rng.integers(0, 3))], errors=(0.02, 0.02, 0.02))
    case.gfa_overlap = int(rng.integers(1, 6))
    path = "/tmp/gfaov_%d.gacase" % it
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
        print("DIFF", it, bad[:4], "overlap", case.gfa_overlap, "rl", rl, flush=True)
    else:
        same += 1
print("gfa overlap %d..%d: identical %d, different %d, reference crashed %d" % (first, first + count - 1, same, differ, crashed))
