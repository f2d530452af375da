"""Every read of bench.py's config-2 batch (10 000 reads, rank 0's set) through the reference: the failed sets must be equal and
every other read identical in score, range, query position, mapping count + checksum, trace-item count + fingerprint.

    python profiles/tools/full_parity.py > profiles/r02_config2_full_parity.txt        (GPU box; ~2 minutes)
"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from graphaligner_b200 import api  # noqa: E402
from graphaligner_b200.tools import gacase  # noqa: E402

case, n = bench.make_workload(2, 0, 1, 1.0, 1)
aligner = api.Aligner(api.Graph.from_case(case))
t0 = time.time()
res = aligner.align(api.PackedReads(case.reads, 10, 0))
path = os.path.join(tempfile.gettempdir(), "full_parity.gacase")
gacase.write_case(case, path)
t1 = time.time()
out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_align"), path, "--quiet", "--summary", "--threads", str(os.cpu_count() or 8)], capture_output=True, text=True)
t2 = time.time()
expected, timing = gacase.parse_ref_output(out.stdout)
r = res.reads
ref_failed = sorted(i for i, e in enumerate(expected) if e["failed"])
my_failed = sorted(int(i) for i in np.nonzero(r["failed"])[0])
ok = [i for i in range(n) if not expected[i]["failed"] and not r["failed"][i]]
mh = gacase.mapping_checksums(r, res.mappings)
diff = {}
for key, col in (("score", "score"), ("start", "alignment_start"), ("end", "alignment_end"), ("qpos", "query_position"), ("nmap", "n_mappings"), ("ntrace", "n_trace")):
    diff[key] = int((r[col][ok].astype(np.int64) != np.array([expected[i][key] for i in ok], dtype=np.int64)).sum())
diff["mapping checksum"] = int((mh[ok] != np.array([expected[i]["mh"] for i in ok], dtype=np.uint64)).sum())
diff["trace fingerprint"] = sum(1 for i in ok if res.trace_hash(i) != expected[i]["th"])
print("config 2 (bench.py's batch: %d reads x 10 kbp, 5 Mbp graph, band 10), CUDA path vs oracle/_ref/ref_align (-O3 -DNDEBUG) on %d threads" % (n, os.cpu_count() or 8))
print("reads failed here:            %s" % [case.reads[i][0] for i in my_failed])
print("reads failed in the reference: %s" % [case.reads[i][0] for i in ref_failed])
print("failed sets equal: %s" % (my_failed == ref_failed))
print("reads compared: %d; differing per field: %s" % (len(ok), diff))
print("reference wall %.1f s (TIME line: %s); CUDA path end to end %.3f s" % (t2 - t1, timing, t1 - t0))
sys.exit(0 if my_failed == ref_failed and not any(diff.values()) else 1)
