"""A/B timing of the general forward kernel (band 35: ga_forward_kernel<S, false>) with the library named in GA_LIB: the device
times of three launches of the same 4 000-read batch (GA_KERNEL_TIMING lines on stderr come from the library itself).
    GA_KERNEL_TIMING=1 [GA_LIB=...] python profiles/tools/general_kernel_ab.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphaligner_b200 import api  # noqa: E402
from graphaligner_b200.tools import synth  # noqa: E402

g = synth.make_graph(77, 400000, chop=32, bubble_every=80, indel_frac=0.3)
case = synth.make_case(77, g, 4000, 3000, b=35, errors=(0.05, 0.05, 0.05))
api.load_library()
aligner = api.Aligner(api.Graph.from_case(case))
for _ in range(3):
    d = aligner.align(case.reads, 35, 0).as_dicts()
print("failed", sum(x["failed"] for x in d), "score sum", sum(x["score"] for x in d if not x["failed"]))
