"""Differential fuzzer (test infrastructure): random graphs and reads (tools.synth), the UNMODIFIED reference hot path
(oracle/_ref/ref_align) as the checker, the CUDA path through the C ABI as the subject.

    python -m graphaligner_b200.tools.fuzz FIRST_SEED COUNT [--keep DIR] [--long-nodes]

--long-nodes: node lengths 300 .. 30 000 bp on backbones of 60 .. 400 kbp (whole-node banding then puts up to ~10^5 columns
into a slice: launch capacities scaled up front, overflow re-runs, the general scratch layout).

Needs a GPU and the oracle built (python __graft_entry__.py).  Prints one line per differing case and a summary; cases the
reference itself crashes on (it segfaults on some cyclic inputs with tiny bands) are counted separately."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from graphaligner_b200.tools import gacase, synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_align")
KEYS = ("failed", "score", "start", "end", "qpos", "nmap", "ntrace", "th")


LONG_NODES = False
RAMP = False


def make_case(it):
    rng = np.random.default_rng(it)
    kind = it % 6
    L = int(rng.integers(2000, 20000))
    chop = int(rng.choice([4, 8, 16, 32, 64, 100]))
    if LONG_NODES:
        L = int(rng.integers(60000, 400000))
        chop = int(rng.choice([300, 1000, 3000, 10000, 30000]))
    kw = dict(chop=chop)
    if kind == 0:
        kw.update(snp_every=int(rng.integers(20, 1000)))
    elif kind == 1:
        kw.update(bubble_every=int(rng.integers(10, 200)), indel_frac=float(rng.random()))
    elif kind == 2:
        kw.update(bubble_every=int(rng.integers(20, 200)), inversion_every=int(rng.integers(200, 2000)))
    elif kind == 3:
        kw.update(bubble_every=int(rng.integers(20, 300)), cycle_every=int(rng.integers(100, 1500)))
    elif kind == 4:
        kw.update(bubble_every=int(rng.integers(50, 300)), tangle_every=int(rng.integers(1000, 5000)), tangle_levels=int(rng.integers(2, 12)),
                  tangle_width=int(rng.integers(2, 5)), tangle_node=int(rng.integers(1, 10)))
    else:
        kw.update(snp_every=int(rng.integers(30, 300)), cycle_every=int(rng.integers(200, 900)), inversion_every=int(rng.integers(300, 1500)))
    g = synth.make_graph(it, L, **kw)
    rl = int(rng.choice([70, 150, 300, 1000, 3000]))
    b = int(rng.choice([2, 5, 10, 20, 35, 50, 100]))
    offs = [(0,), (0, rl // 2, -50), (rl // 3,), (-1,), (1,)][int(rng.integers(0, 5))]
    err = float(rng.choice([0.0, 0.02, 0.05, 0.1]))
    B = 0
    if RAMP:
        # -B ramp: a narrow band that loses noisy reads now and then, a wide backup band (GraphAligner.h:2612-2719)
        b = int(rng.choice([2, 3, 5, 8]))
        B = b + int(rng.choice([5, 15, 40]))
        err = float(rng.choice([0.05, 0.1, 0.15]))
        rl = int(rng.choice([300, 1000, 3000]))
    case = synth.make_case(it, g, 12, rl, b=b, B=B, seed_offsets=offs, decoys=int(rng.integers(0, 2)), errors=(err, err, err), len_jitter=min(rl // 4, 40))
    return case, dict(kind=kind, rl=rl, b=b, B=B, **kw)


def main():
    global LONG_NODES, RAMP
    LONG_NODES = "--long-nodes" in sys.argv
    RAMP = "--ramp" in sys.argv
    redo_same = redo_differ = redo_crashed = reads_redo = reads_redo_differ = reads_stale = cases_stale = stream_errors = 0
    first, count = int(sys.argv[1]), int(sys.argv[2])
    keep = sys.argv[sys.argv.index("--keep") + 1] if "--keep" in sys.argv else None
    from graphaligner_b200 import api
    api.load_library()
    same = differ = ref_crashed = 0
    tmp = tempfile.mkdtemp(prefix="ga_fuzz_")
    for it in range(first, first + count):
        case, desc = make_case(it)
        path = os.path.join(keep or tmp, "fz_%d.gacase" % it)
        gacase.write_case(case, path)
        ref = subprocess.run([REF, path, "--quiet", "--threads", "4"], capture_output=True, text=True)
        if ref.returncode != 0:
            ref_crashed += 1
            if RAMP:
                aligner = api.Aligner(api.Graph.from_case(case))
                res = aligner.align(case.reads, case.b, case.B)
                d = res.as_dicts()
                redo_crashed += 1 if any(m["flags"] & 16 for m in d) else 0
                stream_errors += 1 if any(m["flags"] & 1 for m in d) else 0   # e.g. checkpoints out of order: the reference reads an empty stretch
                res.free()
                aligner.close()
            if not keep:
                os.remove(path)
            continue
        expected, _ = gacase.parse_ref_output(ref.stdout)
        aligner = api.Aligner(api.Graph.from_case(case))
        res = aligner.align(case.reads, case.b, case.B)
        mine = res.as_dicts()
        bad = [e["name"] for m, e in zip(mine, expected)
               if any(m[k] != e[k] for k in KEYS) or [tuple(x) for x in m["mappings"]] != [tuple(x) for x in e["mappings"]]]
        res.free()
        aligner.close()
        if RAMP:
            redo = [m["name"] if "name" in m else i for i, m in enumerate(mine) if m["flags"] & 16]
            reads_redo += len(redo)
            nstale = sum(1 for m in mine if m["flags"] & 32)   # GA_FLAG_RAMP_STALE: the reference's stale checkpoint decided part of the trace
            reads_stale += nstale
            cases_stale += 1 if nstale else 0
            reads_redo_differ += sum(1 for i, (m, e) in enumerate(zip(mine, expected)) if (m["flags"] & 16) and e["name"] in bad)
            if redo:
                if bad:
                    redo_differ += 1
                else:
                    redo_same += 1
                    if keep:
                        print("REDO-IDENTICAL seed %d (%d reads with a redo) %s" % (it, len(redo), desc), flush=True)
        if bad or len(mine) != len(expected):
            differ += 1
            print("DIFF seed %d reads %s %s" % (it, bad[:4], desc), flush=True)
        else:
            same += 1
            if not keep:
                os.remove(path)
    print("fuzz %d..%d: identical %d, different %d, reference crashed %d" % (first, first + count - 1, same, differ, ref_crashed), flush=True)
    if RAMP:
        print("ramp: cases where a redo fired: identical %d, different %d, reference crashed %d; reads with a redo %d, of them different %d"
              % (redo_same, redo_differ, redo_crashed, reads_redo, reads_redo_differ), flush=True)
        print("ramp: reference finished: %d reads in %d cases went through a stale checkpoint (GA_FLAG_RAMP_STALE); reference crashed: %d of those cases report a stream error here"
              % (reads_stale, cases_stale, stream_errors), flush=True)
    return 1 if differ else 0


if __name__ == "__main__":
    sys.exit(main())
