"""SASS of the forward kernel's word step (and an opcode histogram of the whole kernel) from the in-tree build, with source lines:
    python profiles/tools/sass_excerpt.py > profiles/r02_wordstep_sass.txt
(cuobjdump -xelf + nvdisasm -g on graphaligner_b200/build/ga_kernels.cu.o; nothing typed by hand)."""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
obj = os.path.join(ROOT, "graphaligner_b200", "build", "ga_kernels.cu.o")
kernel = sys.argv[1] if len(sys.argv) > 1 else "ga_fast_kernelILi17E"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith("//-----") and ".text." in l and kernel in l][0]
end = [i for i, l in enumerate(sass) if i > start and l.startswith("//-----") and ".text." in l][0]
src = open(os.path.join(ROOT, "graphaligner_b200", "csrc", "ga_fast.cuh")).read().split("\n")
lo = [i for i, l in enumerate(src) if "auto wordStep = [&]" in l][0] + 1
hi = [i for i, l in enumerate(src) if "wordStep(isFirst);" in l][0]
cur = None
ops = Counter()
step = []
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ops[re.sub(r"^@!?U?P\d+\s+", "", m.group(2)).split()[0].split(".")[0]] += 1
        if cur and cur[0] == "ga_fast.cuh" and lo <= cur[1] <= hi:
            step.append("%4d  %s  %s" % (cur[1], m.group(1), m.group(2)))
print("kernel %s: %d SASS instructions; opcodes: %s" % (kernel, sum(ops.values()), ", ".join("%s %d" % kv for kv in ops.most_common(24))))
print("cp.async (LDGSTS) %d, 128-bit global loads/stores: see LDG.E.128 / STG.E.128 below; no tensor-core or TMA opcodes (an integer recurrence)" % ops.get("LDGSTS", 0))
print()
print("word step (ga_fast.cuh:%d-%d, both copies of the two-columns-per-pass loop): %d instructions" % (lo, hi, len(step)))
print("line  addr  instruction")
print("\n".join(step))
