import sys, time, os
sys.path.insert(0, '/root/repo')
os.environ.setdefault("GA_HOST_THREADS", str(os.cpu_count()))
import numpy as np
from graphaligner_b200 import api
from graphaligner_b200.tools import synth
api.load_library()
g, kw = synth.config2(1.0)
case = synth.make_case(1, g, kw['n_reads'], kw['read_len'], b=kw['b'], errors=(0.05, 0.05, 0.05))
graph = api.Graph.from_case(case)
al = api.Aligner(graph)
packed = api.PackedReads(case.reads, 10, 0)
for _ in range(3): al.align(packed).free()
for i in range(6):
    t0 = time.perf_counter(); r = al.align(packed); t1 = time.perf_counter(); x = int(r.reads["score"][0]); t2 = time.perf_counter(); r.free(); t3 = time.perf_counter()
    print("align %.1f  view %.2f  free %.2f  total %.1f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t3-t0)*1e3), flush=True)
