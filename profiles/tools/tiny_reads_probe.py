"""Reads of 1 .. 70 bp (shorter than one 64-row slice), seeds at 0 / middle / last base, against the reference run on the box.
python profiles/tools/tiny_reads_probe.py FIRST COUNT"""
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
    g = synth.make_graph(it, 3000, chop=int(rng.choice([4, 16, 32])), bubble_every=int(rng.integers(20, 200)))
    reads = []
    for k in range(16):
        ln = int(rng.integers(1, 71))
        r = synth.simulate_read(rng, g, max(2, ln), 0.03, 0.03, 0.03)
        if r is None:
            continue
        read, real, walk, mp = r
        read = read[:ln] if ln >= 1 else read
        offs = [0, len(read) // 2, len(read) - 1][int(rng.integers(0, 3))]
        seeds = synth.seeds_for(walk, mp, len(read), [min(offs, len(mp) - 1)])
        reads.append(("r%d" % k, read, seeds))
    case = gacase.Case(list(g.nodes), list(g.edges), reads, int(rng.choice([2, 5, 10])), 0)
    path = "/tmp/tiny_%d.gacase" % it
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
        print("DIFF", it, bad[:4], [(len(r[1]), r[2]) for r in reads if r[0] in bad][:4], flush=True)
    else:
        same += 1
print("tiny reads %d..%d: identical %d, different %d, reference crashed %d" % (first, first + count - 1, same, differ, crashed))
