"""Two contexts on one GPU, one host thread each, alternating batches: does batch i+1's host work (plan, parts, H2D,
assembly, packing) hide behind batch i's kernel?  Prints ms per batch for 1 context and for 2 contexts."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("GA_HOST_THREADS", str(os.cpu_count()))
from graphaligner_b200 import api
from graphaligner_b200.tools import synth

api.load_library()
g, kw = synth.config2(1.0)
case = synth.make_case(1000, g, kw["n_reads"], kw["read_len"], b=kw["b"])
graph = api.Graph.from_case(case)
packed = api.PackedReads(case.reads, 10, 0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
DEPTHS = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 2, 3]
for n_ctx in DEPTHS:
    als = [api.Aligner(graph) for _ in range(n_ctx)]
    for al in als:
        for _ in range(2):
            al.align(packed).free()

    def loop(al, k):
        for _ in range(k):
            r = al.align(packed)
            _ = int(r.reads["score"][0])
            r.free()

    t0 = time.perf_counter()
    ths = [threading.Thread(target=loop, args=(al, K)) for al in als]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    print("S=%s contexts %d: %.1f ms per batch (%d batches)" % (os.environ.get("GA_STREAMS_PER_WARP", "auto"), n_ctx, dt / (K * n_ctx) * 1e3, K * n_ctx), flush=True)
    for al in als:
        al.close()
