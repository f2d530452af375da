"""Stage config 2 (read set replicated R times) once and run the alignment kernel N times with the inputs resident:
python profiles/tools/resident_probe.py R N   (batch-size / streams-per-warp studies and ncu captures of those regimes)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("GA_HOST_THREADS", str(os.cpu_count()))
from graphaligner_b200 import api
from graphaligner_b200.tools import synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
api.load_library()
g, kw = synth.config2(1.0)
case = synth.make_case(1000, g, kw["n_reads"], kw["read_len"], b=kw["b"])
reads = [("%s_c%d" % (n, c), s, sd) for c in range(R) for (n, s, sd) in case.reads]
al = api.Aligner(api.Graph.from_case(case))
packed = api.PackedReads(reads, 10, 0)
staged = al.stage(packed)
al.run(staged)
al.sync()
t0 = time.perf_counter()
for _ in range(N):
    al.run(staged)
al.sync()
ms = (time.perf_counter() - t0) / N * 1e3
res = al.finish(staged, keepalive=packed)
wc = int(res.reads["word_columns"].sum())
print("R=%d streams=%d kernel %.2f ms  %.1f GCUPS  S=%s" % (R, len(reads), ms, wc * 64 / ms / 1e6, os.environ.get("GA_STREAMS_PER_WARP", "auto")), flush=True)
