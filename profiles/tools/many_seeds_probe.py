"""Reads with 4 .. 12 seeds each (true hits along the read plus random decoys): the seed loop's "already aligned" pruning and
best-seed choice (GraphAligner.h:420-450) against the reference run on the box.  python profiles/tools/many_seeds_probe.py FIRST COUNT"""
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
    g = synth.make_graph(it, 20000, chop=int(rng.choice([16, 32])), bubble_every=int(rng.integers(40, 300)), inversion_every=int(rng.choice([0, 1500])))
    rl = int(rng.choice([800, 2500]))
    offs = tuple(sorted(set(int(x) for x in rng.integers(0, rl - 100, int(rng.integers(3, 8))))))
    case = synth.make_case(it, g, 8, rl, b=int(rng.choice([5, 10, 20])), seed_offsets=offs, decoys=int(rng.integers(1, 6)), errors=(0.04, 0.04, 0.04))
    path = "/tmp/seeds_%d.gacase" % it
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
        print("DIFF", it, bad[:4], "seeds", len(case.reads[0][2]), flush=True)
    else:
        same += 1
print("many seeds %d..%d: identical %d, different %d, reference crashed %d" % (first, first + count - 1, same, differ, crashed))
