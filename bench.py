#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's headline workload (configs[1]):
synthetic 5 Mbp linear-ish graph (32-bp nodes, one SNP bubble per 1000 bp), 10 000 SimulateReads-style reads of
10 kbp at ~15 % error, one true seed at read offset 0, band 10.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, one process per GPU)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU aligner on the host cores

A step = one pass of the hot path over one batch (all reads of the rank).  `value` is aligned bp/s with the
inputs already resident in HBM (kernel only, CUDA events on the launching stream); `e2e` is the same metric
through the C ABI with host buffers (read splitting, H2D, kernel, D2H, result assembly inside the timed region):
`e2e.value` streams the K batches through ga_pipeline_* (two contexts per GPU, fill and drain inside the timed region),
`e2e.single_call_ms` is one blocking ga_align_batch per step.  Reads are sharded across ranks with the graph replicated; no data-path collective (weak scaling:
every rank aligns its own 10 000 reads).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRAPH_BP = 5_000_000
READS_PER_GPU = 10_000
READ_LEN = 10_000
BAND = 10
# algorithmic work per forward word update (64 cells), DESIGN.md "Roofline":
BYTES_PER_WORD_COLUMN = 72.25    # 0.25 bases + 4 previous-slice end state in + 4 end state out + 64-byte history record (VP, VN, scores, traceback masks)
LANE_OPS_PER_WORD_COLUMN = 50.0
PIPELINE_DEPTH = int(os.environ.get("GA_PIPELINE_DEPTH", "2"))


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for s in self.samples:
            for i, n in enumerate(names):
                if len(s) > 2 + i and s[2 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(rank, n_reads, scale):
    from graphaligner_b200.tools import synth
    g = synth.make_graph(1, int(GRAPH_BP * scale), chop=32, snp_every=1000)
    case = synth.make_case(1000 + rank, g, n_reads, READ_LEN, b=BAND)
    return g, case


def run_reference_sample(case, sample, threads, tmpdir):
    """The reference's own AlignOneWay (oracle/_ref/ref_align, unmodified sources) on `sample` reads, `threads` workers."""
    from graphaligner_b200.tools import gacase
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_align")
    if not os.path.exists(ref):
        return None
    sub = gacase.Case(case.nodes, case.edges, case.reads[:sample], case.b, case.B, case.gfa_overlap)
    path = os.path.join(tmpdir, "bench_sample.gacase")
    gacase.write_case(sub, path)
    res = subprocess.run([ref, path, "--quiet", "--threads", str(threads)], capture_output=True, text=True)
    if res.returncode != 0:
        return None
    _, timing = gacase.parse_ref_output(res.stdout)
    return timing


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrinks graph and read count (testing only; 1.0 = BASELINE config)")
    ap.add_argument("--cpu-sample", type=int, default=800)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicate", type=int, default=1, help="align R copies of the read set per step (batch-size study only; not the BASELINE config)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_reads = max(32, int(READS_PER_GPU * args.scale))
    cores = os.cpu_count() or 1
    tmpdir = os.environ.get("TMPDIR", "/tmp")
    workload = {"workload": "configs[1]: synthetic %.1f Mbp graph (32-bp nodes, SNP bubble/1000 bp), %d reads x %d bp per GPU, ~15%% error, 1 seed at offset 0, band %d"
                % (GRAPH_BP * args.scale / 1e6, n_reads, READ_LEN, BAND),
                "reads_per_gpu": n_reads, "read_len": READ_LEN, "band": BAND, "graph_bp": int(GRAPH_BP * args.scale),
                "l2": "inputs + DP history (GBs) exceed the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        g, case = make_workload(0, min(n_reads, args.cpu_sample), args.scale)
        sample = min(len(case.reads), args.cpu_sample)
        times, bp = [], 0
        for i in range(args.warmup + args.steps):
            t = run_reference_sample(case, sample, cores, tmpdir)
            if t is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_align missing or failed"}))
                return 0
            if i >= args.warmup:
                times.append(t["wall_ms"])
                bp = t["aligned_bp"]
        ms = sum(times) / len(times)
        val = bp / (ms * 1e-3)
        line = {"impl": "reference", "metric": "aligned_bp_per_s", "value": val, "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload,
                "cpu_baseline": {"value": val, "unit": "bp/s", "cores": cores, "kind": "reference", "sample": "%d reads of the workload per step" % sample},
                "e2e": {"value": val, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # stdout carries the JSON line and nothing else: library chatter on fd 1 (NCCL prints its version there) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    # host worker threads per rank: the ranks of one box share its cores
    os.environ.setdefault("GA_HOST_THREADS", str(max(1, cores // max(1, world))))
    import torch
    import torch.distributed as dist
    from graphaligner_b200 import api, multi_gpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    g, case = make_workload(rank, n_reads, args.scale)
    graph = api.Graph.from_case(case)
    aligner = api.Aligner(graph, device=local_rank)
    if args.replicate > 1:
        case.reads = [("%s_c%d" % (n, c), s_, sd) for c in range(args.replicate) for (n, s_, sd) in case.reads]
        workload["workload"] += " x%d replicated (batch-size study)" % args.replicate
    packed = api.PackedReads(case.reads, case.b, case.B)
    total_bp = packed.total_bp

    stream = torch.cuda.ExternalStream(aligner.cuda_stream(), device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident: inputs staged once, K kernel-only steps -------------------------------------------------------
    staged = aligner.stage(packed)
    for _ in range(args.warmup):
        aligner.run(staged)
    aligner.sync()
    aligner.reset_stats()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        for a, b in evs:
            a.record(stream)
            aligner.run(staged)
            b.record(stream)
        aligner.sync()
        barrier()
        wall = time.perf_counter() - t0
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    res = aligner.finish(staged, keepalive=packed)
    aligned_bp = int(sum(len(case.reads[i][1]) for i in range(len(case.reads)) if not res.reads["failed"][i]))
    word_columns = int(res.reads["word_columns"].sum())
    failed = int(res.reads["failed"].sum())
    stats = aligner.stats()
    res.free()
    aligner.free_staged(staged)
    int_peak = aligner.int32_peak() if rank == 0 else 0.0
    # device time of the K steps: MAX over ranks; work: SUM over ranks (no data-path collective anywhere else)
    d = dist if world > 1 else None
    dev_ms = multi_gpu.reduce_max(d, sum(kernel_ms), device="cuda")
    aligned_bp_all, word_columns_all, failed_all = multi_gpu.reduce_sum(d, [aligned_bp, word_columns, failed], device="cuda")
    ms_per_step = dev_ms / args.steps
    value = aligned_bp_all / (ms_per_step * 1e-3)
    gcups = word_columns_all * 64 / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers --------------------------------------------------------
    # (1) one blocking ga_align_batch call per step: the latency of a single batch
    for _ in range(min(2, args.warmup)):
        aligner.align(packed).free()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = aligner.align(packed)
        _ = int(r.reads["score"][0])
        r.free()
    torch.cuda.synchronize()
    single_s = multi_gpu.reduce_max(d, time.perf_counter() - t0, device="cuda")
    aligner.close()
    # (2) the same K batches as a stream through ga_pipeline_* (two contexts on this GPU, one host thread each): every
    # step still pads and uploads its reads from host memory and brings its results back to the host; step i+1's
    # host work runs while step i's kernel does.  This is the throughput number (`e2e.value`).
    pipe = api.Pipeline(graph, device=local_rank, depth=PIPELINE_DEPTH)
    # warm-up: every lane's grow-only device pools reach their size, and the results are held until the end so that the
    # process-wide pools of pinned / pageable result blocks hold more blocks than can be alive at once in the timed loop
    # (a miss there is a cudaHostAlloc of ~120 MB: 40-60 ms in the middle of a 5-step measurement)
    held = list(pipe.align_all([packed] * max(2 * PIPELINE_DEPTH, args.warmup)))
    for r in held:
        r.free()
    del held
    pipe.reset_stats()
    barrier()
    t0 = time.perf_counter()
    checksum = 0
    for r in pipe.align_all([packed] * args.steps):
        checksum += int(r.reads["score"][0])
        r.free()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    st = pipe.stats()
    pipe.close()
    e2e_s = multi_gpu.reduce_max(d, e2e_s, device="cuda")
    e2e_val = aligned_bp_all * args.steps / e2e_s

    line = None
    if rank == 0:
        hbm_peak, peak_kind = load_peaks()
        mean_kernel_s = (sum(kernel_ms) / len(kernel_ms)) * 1e-3
        alg_bytes = word_columns * BYTES_PER_WORD_COLUMN
        achieved = alg_bytes / mean_kernel_s / 1e9
        lane_ops = word_columns * LANE_OPS_PER_WORD_COLUMN
        line = {"metric": "aligned_bp_per_s", "value": value, "unit": "bp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": workload, "gcups": gcups, "word_columns_per_step": word_columns_all, "failed_reads": failed_all,
                "clocks": clocks.summary(),
                "e2e": {"value": e2e_val, "unit": "bp/s", "h2d_bytes_per_step": st["h2d_bytes"] // args.steps, "d2h_bytes_per_step": st["d2h_bytes"] // args.steps,
                        "ms_per_step": e2e_s / args.steps * 1e3, "mode": "ga_pipeline, depth %d (K batches streamed, results in order)" % PIPELINE_DEPTH,
                        "single_call_ms": single_s / args.steps * 1e3},
                # kernels of ours launched inside the timed region: ga_peq_kernel + ga_align_kernel per step
                "gpu_launches": int(st["launches"]) if st["launches"] else 2 * args.steps,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": load_traffic(),
                             "peak_source": peak_kind, "kernel": "ga_align_kernel", "kernel_ms": mean_kernel_s * 1e3,
                             "int_alu": {"achieved_lane_ops_per_s": lane_ops / mean_kernel_s, "peak_lane_ops_per_s": int_peak,
                                         "frac": (lane_ops / mean_kernel_s / int_peak) if int_peak else None,
                                         "note": "50 INT32 lane-ops per forward word update (SURVEY 8d); peak = measured LOP3/IADD3 probe on this GPU"}},
                "wall_s_resident": wall}
        if world == 1 and not args.no_cpu_baseline:
            sample = min(len(case.reads), args.cpu_sample)
            t = run_reference_sample(case, sample, cores, tmpdir)
            if t is not None:
                line["cpu_baseline"] = {"value": t["aligned_bp"] / (t["wall_ms"] * 1e-3), "unit": "bp/s", "cores": cores, "kind": "reference",
                                        "sample": "first %d reads of the workload, reference AlignOneWay (-O3 -DNDEBUG), %d threads" % (sample, cores)}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "bp/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref/ref_align unavailable"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
