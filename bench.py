#!/usr/bin/env python
"""Benchmark of the hot path (seeded AlignOneWay: banded bit-parallel DP of reads against the graph + traceback) on the
BASELINE.json workloads.  Default = the headline workload configs[1]:
synthetic 5 Mbp linear-ish graph (32-bp nodes, one SNP bubble per 1000 bp), 10 000 SimulateReads-style reads of
10 kbp at ~15 % error, one true seed at read offset 0, band 10.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, one process per GPU)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU aligner on the host cores
  python bench.py --config {3,4,5} [--scale S]                   the other BASELINE workloads, each with a
                                                                 parity sample against the reference inside the run

A step = one pass of the hot path over the rank's read set (config 2: one batch of 10 000 reads; the larger configs: their
reads in batches of `batch_reads`).  `value` is aligned bp/s with the inputs already resident in HBM (the three kernels of
a launch sequence only, CUDA events on the launching stream; for the multi-batch configs measured on the first batch);
`e2e` is the same metric through the C ABI with host buffers (read splitting, H2D, kernels, D2H, result assembly inside
the timed region): `e2e.value` streams K passes over all batches through ga_pipeline_* (two contexts per GPU, fill and
drain inside the timed region), `e2e.single_call_ms` is one blocking ga_align_batch per step (config 2).  Reads are
sharded across ranks with the graph replicated; no data-path collective (weak scaling: every rank aligns its own reads).
Prints ONE JSON line on rank 0.
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

# algorithmic work per forward word update (64 cells), SURVEY.md 8d:
LANE_OPS_PER_WORD_COLUMN = 50.0          # INT32 lane-operations (+150 per extra in-edge merge, not counted here)
BYTES_PER_WORD_COLUMN_CHECKPOINT = 8.25  # 0.25 B bases + 4 B previous end state in + 4 B end state out
BYTES_PER_WORD_COLUMN_STORE = 24.25      # + 16 B {VP, VN} kept for the traceback (what this implementation does, plus a 4 B score word)
PIPELINE_DEPTH = int(os.environ.get("GA_PIPELINE_DEPTH", "2"))

# reads per GPU at scale 1.0; graph_bp = backbone length; batch = reads per ga_batch; resident = reads of the batch the
# kernel-only arm is timed on; parity = reads of the in-run sample against the reference; cpu = reads of the cpu_baseline sample
CONFIGS = {
    2: dict(name="configs[1]", graph_bp=5_000_000, reads=10_000, read_len=10_000, bands=[10], batch=10_000, resident=10_000, parity=0, cpu=800,
            what="synthetic %.1f Mbp graph (32-bp nodes, SNP bubble/1000 bp), %d reads x %d bp per GPU, ~15%% error, 1 seed at offset 0"),
    3: dict(name="configs[2]", graph_bp=100_000_000, reads=100_000, read_len=10_000, bands=[10], batch=10_000, resident=4_000, parity=64, cpu=64,
            what="synthetic %.1f Mbp variation graph (32-bp nodes, a bubble per 100 bp: 80%% SNP / 20%% indel, inversions), %d reads x %d bp per GPU, "
                 "~15%% error, PickSeedHits-style seeds at read offsets 0 / 5000 / end-300 + 1 decoy"),
    4: dict(name="configs[3]", graph_bp=3_000_000_000, reads=125_000, read_len=15_000, bands=[10], batch=10_000, resident=10_000, parity=32, cpu=32,
            what="synthetic %.1f Mbp GFA graph (0M overlaps, 32-bp nodes, SNP bubble/1000 bp + indel bubble/5000 bp), replicated per GPU, "
                 "%d reads x %d bp per GPU (sharded contiguously by index), ~15%% error, 1 seed at offset 0"),
    5: dict(name="configs[4]", graph_bp=100_000_000, reads=2_000, read_len=50_000, bands=[5, 10, 20, 35, 50, 75, 100], batch=2_000, resident=2_000, parity=16, cpu=16,
            what="config-3 style %.1f Mbp graph + a tangle per 250 kbp (20 levels x 4 parallel 8-bp nodes, back-edges forming 2-node cycles), "
                 "%d reads x %d bp per GPU, ~15%% error, 1 seed at offset 0, band sweep"),
}


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/r02_traffic.json,
    written by profiles/tools/ncu_export.py from the raw CSV next to it), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            return d.get("dram_bytes_per_launch"), d.get("kernel"), name
    return None, None, None


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


def make_workload(config, rank, world, scale, workers, only_reads=None, reads_override=None):
    """Graph (the same on every rank) and this rank's reads.  Config 2 keeps the generator calls of round 1 (same inputs);
    the larger configs build graph and reads on `workers` processes."""
    from graphaligner_b200.tools import synth
    cfg = CONFIGS[config]
    n_reads = max(32, int(cfg["reads"] * scale)) if reads_override is None else reads_override
    if only_reads is not None:
        n_reads = min(n_reads, only_reads)
    gfa_overlap = None
    if config == 2:
        g = synth.make_graph(1, int(cfg["graph_bp"] * scale), chop=32, snp_every=1000)
        case = synth.make_case(1000 + rank, g, n_reads, cfg["read_len"], b=10)
        return case, n_reads
    par = workers > 1
    if config == 3:
        g, kw = synth.config3(scale, parallel=par)
        extra = dict(seed_offsets=kw["seed_offsets"], decoys=kw["decoys"])
    elif config == 4:
        g, kw = synth.config4(scale, parallel=par)
        extra = {}
        gfa_overlap = kw["gfa_overlap"]
    else:
        g, kw = synth.config5(scale, parallel=par)
        extra = {}
    # every rank simulates its own shard: chunk seeds depend on the global chunk index, so the shards are the pieces of ONE read set
    chunk = 250 if config == 5 else 500
    per = -(-n_reads // chunk) * chunk
    case = synth.make_case_parallel(config, g, per * world, cfg["read_len"], workers=workers, chunk=chunk, read_range=(rank * per, rank * per + n_reads),
                                    b=10, errors=(0.05, 0.05, 0.05), **extra)
    case.gfa_overlap = gfa_overlap
    return case, n_reads


def run_reference(case, reads, band, threads, tmpdir, flavour="ref_align", tag="bench"):
    """The reference's own AlignOneWay (oracle/_ref/<flavour>, unmodified sources) on `reads`, `threads` workers.
    Returns (per-read result dicts, timing) or (None, None)."""
    from graphaligner_b200.tools import gacase
    ref = os.path.join(ROOT, "oracle", "_ref", flavour)
    if not os.path.exists(ref):
        return None, None
    sub = gacase.Case(case.nodes, case.edges, reads, band, 0, case.gfa_overlap)
    path = os.path.join(tmpdir, "%s_sample_%d.gacase" % (tag, os.getpid()))
    gacase.write_case(sub, path)
    res = subprocess.run([ref, path, "--quiet", "--threads", str(threads)], capture_output=True, text=True)
    try:
        os.unlink(path)
    except OSError:
        pass
    if res.returncode != 0:
        return None, None
    return gacase.parse_ref_output(res.stdout)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json workload, 1-based (2 = configs[1], the headline)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrinks graph and read count (1.0 = BASELINE size; config 4 at 1.0 needs a native generator)")
    ap.add_argument("--reads", type=int, default=None, help="reads per GPU instead of the config's (with --scale: the graph alone is scaled)")
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--parity-sample", type=int, default=None)
    ap.add_argument("--bands", type=str, default=None, help="comma-separated band widths instead of the config's")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stock-cpu", action="store_true", help="skip the stock-flags (asserts on, reference makefile:3) flavour of the reference on the cpu sample")
    ap.add_argument("--stock-cpu", action="store_true", help="(default now; kept for old command lines)")
    ap.add_argument("--replicate", type=int, default=1, help="align R copies of the read set per step (batch-size study only; not the BASELINE config)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = dict(CONFIGS[args.config])
    if args.bands:
        cfg["bands"] = [int(x) for x in args.bands.split(",")]
    if args.cpu_sample is not None:
        cfg["cpu"] = args.cpu_sample
    if args.parity_sample is not None:
        cfg["parity"] = args.parity_sample
    cores = os.cpu_count() or 1
    workers = max(1, cores // max(1, world))
    tmpdir = os.environ.get("TMPDIR", "/tmp")
    band0 = 10 if 10 in cfg["bands"] else cfg["bands"][0]

    def workload_desc(n_reads):
        return {"workload": "%s: " % cfg["name"] + cfg["what"] % (cfg["graph_bp"] * args.scale / 1e6, n_reads, cfg["read_len"]) + ", band %s" % "/".join(str(b) for b in cfg["bands"]),
                "reads_per_gpu": n_reads, "read_len": cfg["read_len"], "band": band0, "graph_bp": int(cfg["graph_bp"] * args.scale),
                "l2": "inputs + DP history (GBs) exceed the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = cfg["cpu"]
        case, n = make_workload(args.config, 0, 1, args.scale, cores, only_reads=sample, reads_override=args.reads)
        reads = case.reads[:sample]
        times, bp = [], 0
        for i in range(args.warmup + args.steps):
            _, t = run_reference(case, reads, band0, cores, tmpdir)
            if t is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_align missing or failed"}))
                return 0
            if i >= args.warmup:
                times.append(t["wall_ms"])
                bp = t["aligned_bp"]
        ms = sum(times) / len(times)
        val = bp / (ms * 1e-3)
        n_full = max(32, int(cfg["reads"] * args.scale)) if args.reads is None else args.reads
        line = {"impl": "reference", "metric": "aligned_bp_per_s", "value": val, "unit": "bp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": workload_desc(n_full),
                "cpu_baseline": {"value": val, "unit": "bp/s", "cores": cores, "kind": "reference", "sample": "%d reads of the workload per step, reference AlignOneWay (-O3 -DNDEBUG)" % len(reads)},
                "e2e": {"value": val, "unit": "bp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # stdout carries the JSON line and nothing else: library chatter on fd 1 (NCCL prints its version there) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    # host worker threads per rank: the ranks of one box share its cores
    os.environ.setdefault("GA_HOST_THREADS", str(workers))
    import torch
    import torch.distributed as dist
    from graphaligner_b200 import api, multi_gpu
    from graphaligner_b200.tools import gacase

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    d = dist if world > 1 else None

    t_prep = time.perf_counter()
    case, n_reads = make_workload(args.config, rank, world, args.scale, workers, reads_override=args.reads)
    workload = workload_desc(n_reads)
    graph = api.Graph.from_case(case)
    aligner = api.Aligner(graph, device=local_rank)
    reads = case.reads
    if args.replicate > 1:
        reads = [("%s_c%d" % (n, c), s_, sd) for c in range(args.replicate) for (n, s_, sd) in reads]
        workload["workload"] += " x%d replicated (batch-size study)" % args.replicate
    nb = cfg["batch"] * args.replicate
    chunks = [reads[i:i + nb] for i in range(0, len(reads), nb)]
    workload["batches_per_step"] = len(chunks)
    t_prep = time.perf_counter() - t_prep
    stream = torch.cuda.ExternalStream(aligner.cuda_stream(), device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    held_pinned = []

    def pinned_alloc(nbytes):
        if os.environ.get("GA_BENCH_PAGEABLE"):
            import numpy as np
            return np.empty(nbytes, dtype=np.uint8)
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        held_pinned.append(t)
        return t.numpy()

    def measure(band, steps, warmup, with_single_call):
        """One band width: the kernel-only arm on the first batch, then the end-to-end arm over all batches."""
        # the read bytes of every batch lie in page-locked host memory (the contract's "inputs from pinned host memory"): the library
        # uploads such a buffer in place; GA_BENCH_PAGEABLE=1 puts them in ordinary memory (then staged through the context's pinned buffer)
        batches = [api.PackedReads(c, band, 0, seq_alloc=pinned_alloc) for c in chunks]
        total_bp = sum(p.total_bp for p in batches)
        # ---- resident: inputs staged once, K kernel-only steps ---------------------------------------------------
        n_res = min(len(chunks[0]), cfg["resident"] * args.replicate)
        staged = packed0 = None
        while staged is None:
            packed0 = batches[0] if n_res == len(chunks[0]) else api.PackedReads(chunks[0][:n_res], band, 0, seq_alloc=pinned_alloc)
            try:
                staged = aligner.stage(packed0)
            except RuntimeError:
                if n_res <= 64:
                    raise
                n_res //= 2   # the batch's DP history does not fit one launch: a smaller resident batch
        for _ in range(warmup):
            aligner.run(staged)
        aligner.sync()
        aligner.reset_stats()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
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
        res = aligner.finish(staged, keepalive=packed0)
        rr = res.reads
        res_bp = int(sum(len(chunks[0][i][1]) for i in range(n_res) if not rr["failed"][i]))
        res_wc = int(rr["word_columns"].sum())
        res_failed = int(rr["failed"].sum())
        res_streams_err = int((rr["flags"] & 1).sum())
        stats = aligner.stats()   # the last launch sequence's kernels (event timers are read in finish)
        res.free()
        aligner.free_staged(staged)
        # device time of the K steps: MAX over ranks; work: SUM over ranks (no data-path collective anywhere else)
        dev_ms = multi_gpu.reduce_max(d, sum(kernel_ms), device="cuda")
        bp_all, wc_all, failed_all = multi_gpu.reduce_sum(d, [res_bp, res_wc, res_failed], device="cuda")
        ms_per_step = dev_ms / steps
        out = {"band": band, "value": bp_all / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "gcups": wc_all * 64 / (ms_per_step * 1e-3) / 1e9,
               "word_columns_per_step": wc_all, "failed_reads": failed_all, "resident_reads": n_res, "stream_errors": res_streams_err,
               "streams": int(stats["streams"]), "streams_rerun_with_general_layout": int(stats["retries"]),
               "kernel_split_ms": {"peq": stats["peq_us"] / 1e3, "forward": stats["forward_us"] / 1e3, "trace": stats["trace_us"] / 1e3},
               "mean_kernel_ms": sum(kernel_ms) / len(kernel_ms), "word_columns_rank0": res_wc,
               # the staged batch's launch counter is read out when it is finished: it also holds the warm-up runs
               "kernel_launches": int(stats["launches"]) * steps // max(1, steps + warmup),
               "clocks": clocks.summary(), "wall_s_resident": wall}
        # ---- end to end through the C ABI with host buffers ----------------------------------------------------
        if with_single_call:
            # one blocking ga_align_batch call per batch: the latency of a single batch
            for _ in range(min(2, warmup)):
                aligner.align(batches[0]).free()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                for p in batches:
                    r = aligner.align(p)
                    _ = int(r.reads["score"][0])
                    r.free()
            torch.cuda.synchronize()
            out["single_call_ms"] = multi_gpu.reduce_max(d, time.perf_counter() - t0, device="cuda") / steps * 1e3
        # the same K passes as a stream through ga_pipeline_* (two contexts on this GPU, one host thread each): every batch
        # still pads and uploads its reads from host memory and brings its results back to the host; batch i+1's host work
        # runs while batch i's kernels do.  This is the throughput number (`e2e.value`).
        pipe = api.Pipeline(graph, device=local_rank, depth=PIPELINE_DEPTH)
        # warm-up: every lane's grow-only device pools reach their size, and the results are held until the end so that the
        # process-wide pools of pinned / pageable result blocks hold more blocks than can be alive at once in the timed loop
        # (a miss there is a cudaHostAlloc of ~100 MB: 40-60 ms in the middle of a measurement)
        n_warm = max(2 * PIPELINE_DEPTH, warmup if len(batches) == 1 else len(batches))
        held = list(pipe.align_all([batches[i % len(batches)] for i in range(n_warm)]))
        for r in held:
            r.free()
        del held
        pipe.reset_stats()
        barrier()
        t0 = time.perf_counter()
        checksum = 0
        aligned = 0
        summaries = []
        arrivals = []
        for k, r in enumerate(pipe.align_all(batches * steps)):
            arrivals.append(time.perf_counter() - t0)
            checksum += int(r.reads["score"][0])
            if k < len(batches):
                ok = r.reads["failed"] == 0
                aligned += int(sum(len(chunks[k][i][1]) for i in range(len(chunks[k])) if ok[i]))
                if cfg["parity"] and rank == 0 and band in parity_bands:
                    summaries.append(r.reads[["failed", "score", "alignment_start", "alignment_end", "n_mappings"]].copy())   # first pass: checked by the parity sample below
            r.free()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        st = pipe.stats()
        e2e_s = multi_gpu.reduce_max(d, e2e_s, device="cuda")
        (aligned_all,) = multi_gpu.reduce_sum(d, [aligned], device="cuda")
        out["e2e"] = {"value": aligned_all * steps / e2e_s, "unit": "bp/s", "h2d_bytes_per_step": st["h2d_bytes"] // steps, "d2h_bytes_per_step": st["d2h_bytes"] // steps,
                      "ms_per_step": e2e_s / steps * 1e3, "mode": "ga_pipeline, depth %d (K passes over %d batch(es) streamed, results in order)" % (PIPELINE_DEPTH, len(batches)),
                      "device_ms_per_step": {"peq": st["peq_us"] / 1e3 / steps, "forward": st["forward_us"] / 1e3 / steps, "trace": st["trace_us"] / 1e3 / steps},
                      "word_columns_per_step": st["word_columns"] // steps, "total_bp_per_step": total_bp,
                      # when each batch's results arrived (ms since the start of the timed region, this rank): fill, steady state, drain
                      "batch_arrival_ms": [round(a * 1e3, 2) for a in arrivals[:32]]}
        out["e2e_launches"] = int(st["launches"])
        pipe.close()
        out["_summaries"] = summaries
        return out

    parity_bands = {band0, cfg["bands"][-1]}   # each parity sample writes the whole graph out for the reference: the headline band and the widest one
    sweep = []
    parity = None
    cpu_line = None
    for band in cfg["bands"]:
        m = measure(band, args.steps, args.warmup, with_single_call=(args.config == 2))
        summaries = m.pop("_summaries")
        if rank == 0 and cfg["parity"] and summaries:
            # ---- parity sample inside the run: every n-th read of rank 0's shard through the reference (same inputs), against (a) the
            # same reads aligned again as one small batch (every mapping, trace fingerprint) and (b) their results in the timed pass ----
            n_s = min(cfg["parity"], len(reads))
            idx = sorted(set(int(i * len(reads) / n_s) for i in range(n_s)))
            sample = [reads[i] for i in idx]
            t0 = time.perf_counter()
            expected, timing = run_reference(case, sample, band, cores, tmpdir, tag="parity")
            if expected is None:
                pr = {"band": band, "sample": len(idx), "status": "reference unavailable or crashed on the sample"}
            else:
                ref_wall = time.perf_counter() - t0
                res = aligner.align(sample, band, 0)
                mine = res.as_dicts()
                same = diff = ref_failed = timed_diff = 0
                first_diff = None
                for i, mn, e in zip(idx, mine, expected):
                    k, j = divmod(i, nb)
                    t = summaries[k][j]
                    ref_failed += 1 if e["failed"] else 0
                    if (int(t["failed"]) != e["failed"]) or (not e["failed"] and (int(t["score"]) != e["score"] or int(t["alignment_start"]) != e["start"]
                                                                                   or int(t["alignment_end"]) != e["end"] or int(t["n_mappings"]) != e["nmap"])):
                        timed_diff += 1
                    if gacase.same_result(mn, e):
                        same += 1
                    else:
                        diff += 1
                        first_diff = first_diff or reads[i][0]
                res.free()
                pr = {"band": band, "sample": len(idx), "identical": same, "different": diff, "timed_pass_summary_different": timed_diff, "reference_failed": ref_failed,
                      "first_different": first_diff, "reference_wall_s": ref_wall}
                if band == band0 and timing is not None and not args.no_cpu_baseline:
                    cpu_line = {"value": timing["aligned_bp"] / (timing["wall_ms"] * 1e-3), "unit": "bp/s", "cores": cores, "kind": "reference",
                                "sample": "%d reads of the workload (every %d-th of rank 0's shard), reference AlignOneWay (-O3 -DNDEBUG), %d threads" % (len(idx), max(1, len(reads) // len(idx)), cores)}
            parity = (parity or []) + [pr]
        sweep.append(m)
    int_peak = aligner.int32_peak() if rank == 0 else 0.0
    aligner.close()

    if rank == 0:
        head = next(m for m in sweep if m["band"] == band0)
        hbm_peak, peak_kind = load_peaks()
        traffic, traffic_kernel, traffic_file = load_traffic()
        # the dominant kernel = the forward DP kernel (ga_fast_kernel / ga_forward_kernel): its own launch duration (events on
        # the context's stream, ga_stats.forward_us) and the word updates it processed
        fwd_s = head["kernel_split_ms"]["forward"] * 1e-3   # the last timed step's launch
        wc = head["word_columns_rank0"]
        lane_ops = wc * LANE_OPS_PER_WORD_COLUMN
        achieved = lane_ops / fwd_s if fwd_s > 0 else 0.0
        line = {"metric": "aligned_bp_per_s", "value": head["value"], "unit": "bp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": workload, "gcups": head["gcups"], "word_columns_per_step": head["word_columns_per_step"], "failed_reads": head["failed_reads"],
                "value_note": "kernel-only (peq + forward + trace of a launch sequence), inputs resident in HBM, %d reads per GPU per step; SURVEY 8d's wall-time metric is e2e" % head["resident_reads"],
                "clocks": head["clocks"],
                "e2e": dict(head["e2e"], **({"single_call_ms": head["single_call_ms"]} if "single_call_ms" in head else {})),
                # kernels of ours launched inside the resident timed region: bad-character check + match masks + forward DP + traceback per launch sequence
                "gpu_launches": head["kernel_launches"],
                "kernel_split_ms": head["kernel_split_ms"],
                "roofline": {"bound": "int_alu", "achieved": achieved / 1e12, "peak": int_peak / 1e12, "unit": "T lane-op/s", "frac": (achieved / int_peak) if int_peak else None,
                             "kernel": "forward DP kernel (ga_fast_kernel<S> for small bands - config 2: S = 17 - else ga_forward_kernel<S>)", "kernel_ms": fwd_s * 1e3,
                             "note": "50 INT32 lane-ops per forward word update (SURVEY 8d) x word updates of one launch / the kernel's own launch duration; "
                                     "peak = dependency-free LOP3/IADD3 probe measured in this run (the contract's hbm|tensor bounds do not bind an integer kernel)",
                             "traffic": traffic, "traffic_source": traffic_file, "traffic_kernel": traffic_kernel,
                             "hbm": {"peak": hbm_peak, "unit": "GB/s", "peak_source": peak_kind,
                                     "achieved_store_mode": wc * BYTES_PER_WORD_COLUMN_STORE / fwd_s / 1e9 if fwd_s > 0 else None,
                                     "frac_store_mode": wc * BYTES_PER_WORD_COLUMN_STORE / fwd_s / 1e9 / hbm_peak if fwd_s > 0 else None,
                                     "frac_checkpoint_mode": wc * BYTES_PER_WORD_COLUMN_CHECKPOINT / fwd_s / 1e9 / hbm_peak if fwd_s > 0 else None,
                                     "algorithmic_bytes_store_mode": wc * BYTES_PER_WORD_COLUMN_STORE,
                                     "traffic_over_algorithmic": (traffic / (wc * BYTES_PER_WORD_COLUMN_STORE)) if (traffic and args.config == 2 and wc) else None}},
                "wall_s_resident": head["wall_s_resident"], "prep_s": t_prep}
        if len(sweep) > 1:
            line["sweep"] = [{k: v for k, v in m.items() if k not in ("clocks", "wall_s_resident")} for m in sweep]
        if parity:
            line["parity"] = parity
        if world == 1 and not args.no_cpu_baseline:
            if cpu_line is None:
                sample = min(len(reads), cfg["cpu"])
                _, t = run_reference(case, reads[:sample], band0, cores, tmpdir)
                if t is not None:
                    cpu_line = {"value": t["aligned_bp"] / (t["wall_ms"] * 1e-3), "unit": "bp/s", "cores": cores, "kind": "reference",
                                "sample": "first %d reads of the workload, reference AlignOneWay (-O3 -DNDEBUG), %d threads" % (sample, cores)}
                else:
                    cpu_line = {"value": None, "unit": "bp/s", "cores": cores, "kind": "reference", "sample": "oracle/_ref/ref_align unavailable"}
            line["cpu_baseline"] = cpu_line
            if not args.no_stock_cpu and args.config == 2:
                sample = min(len(reads), cfg["cpu"])
                _, t = run_reference(case, reads[:sample], band0, cores, tmpdir, flavour="ref_align_stock")
                line["cpu_baseline_stock"] = ({"value": t["aligned_bp"] / (t["wall_ms"] * 1e-3), "unit": "bp/s", "cores": cores, "kind": "reference",
                                               "sample": "first %d reads, stock flags (-O3 -g, asserts on: reads that assert count as failed), %d threads" % (sample, cores)}
                                              if t is not None else {"value": None, "sample": "ref_align_stock unavailable or crashed"})
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
