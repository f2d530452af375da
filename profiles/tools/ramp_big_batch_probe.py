"""-B ramp at batch scale: one case of N noisy reads with a narrow band and a wide backup band, aligned by the library named in GA_LIB
(default: the CUDA library) and dumped as one line per read.  Run once on the B200 and once through the CPU emulation
(GA_LIB=oracle/_ref/libga_hostsim.so, one stream per "warp") and diff the two files: the RAMP instantiation of the general kernel runs
32 streams per warp in lock step with lanes in different phases (forward pass / re-computed stretches), the emulation one at a time.
    python profiles/tools/ramp_big_batch_probe.py OUT.txt [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphaligner_b200 import api  # noqa: E402
from graphaligner_b200.tools import synth  # noqa: E402

n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
g = synth.make_graph(4242, 300000, chop=32, bubble_every=90, indel_frac=0.4)
case = synth.make_case(4242, g, n, 2000, b=2, B=22, errors=(0.13, 0.13, 0.13), len_jitter=300)
api.load_library()
aligner = api.Aligner(api.Graph.from_case(case))
d = aligner.align(case.reads, case.b, case.B).as_dicts()
with open(sys.argv[1], "w") as f:
    for x in d:
        f.write("%s failed=%d flags=%d score=%d start=%d end=%d qpos=%d nmap=%d ntrace=%d th=%s\n" % (x["name"], x["failed"], x["flags"] & ~8, x["score"], x["start"], x["end"], x["qpos"], x["nmap"], x["ntrace"], x["th"]))
print("reads %d, with a redo %d, through a stale checkpoint %d, failed %d, stream errors %d" % (len(d), sum(1 for x in d if x["flags"] & 16), sum(1 for x in d if x["flags"] & 32),
      sum(x["failed"] for x in d), sum(1 for x in d if x["flags"] & 1)))
print(aligner.stats())
