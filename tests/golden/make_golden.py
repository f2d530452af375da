"""Regenerates the golden fixtures: seeded synthetic cases (graphaligner_b200.tools.synth) and the
reference's own test/smallexample, each run through the UNMODIFIED reference hot path
(oracle/_ref/ref_align, built by oracle/Makefile from /root/reference) -> <name>.expected.
Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphaligner_b200.tools import gacase, synth, vgio  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_align")


def smallexample():
    base = "/root/reference/test/smallexample/"
    n, e = vgio.load_vg_graph(base + "sub_test.vg")
    seeds = vgio.load_gam(base + "seedalignment.gam")
    reads = vgio.load_fastq(base + "read.fastq")
    by = {}
    for s in seeds:
        pos = s["path"][0]["position"]
        by.setdefault(s["name"], []).append((pos["node_id"], s["query_position"], pos["is_reverse"]))
    return gacase.Case([(x["id"], x["sequence"]) for x in n], [(x["from"], x["from_start"], x["to"], x["to_end"]) for x in e],
                       [(name, seq, by.get(name, [])) for name, seq in reads], 10, 0)


def fuzz_case(it):
    """The differential fuzzer's recipe for seed `it` (kind 3: bubbles + short cycles)."""
    import numpy as np
    rng = np.random.default_rng(it)
    assert it % 6 == 3
    L = int(rng.integers(2000, 20000))
    chop = int(rng.choice([4, 8, 16, 32, 64, 100]))
    g = synth.make_graph(it, L, chop=chop, bubble_every=int(rng.integers(20, 300)), cycle_every=int(rng.integers(100, 1500)))
    rl = int(rng.choice([70, 150, 300, 1000, 3000]))
    b = int(rng.choice([2, 5, 10, 20, 35, 50, 100]))
    offs = [(0,), (0, rl // 2, -50), (rl // 3,), (-1,), (1,)][int(rng.integers(0, 5))]
    err = float(rng.choice([0.0, 0.02, 0.05, 0.1]))
    return synth.make_case(it, g, 12, rl, b=b, seed_offsets=offs, decoys=int(rng.integers(0, 2)), errors=(err, err, err), len_jitter=min(rl // 4, 40))


def cases():
    yield "smallexample", smallexample()
    # config-2 style: chopped backbone, SNP bubble per 1000 bp, seed at read offset 0
    yield "dag_snp", synth.make_case(11, synth.make_graph(11, 40000, chop=32, snp_every=1000), 16, 3000, b=10)
    # config-3 style: dense bubbles (SNP + indel), inversions, 3 seeds + decoy per read
    yield "bubbles_multiseed", synth.make_case(12, synth.make_graph(12, 30000, chop=32, bubble_every=100, inversion_every=2000), 12, 2500, b=10,
                                               seed_offsets=(0, 1200, -300), decoys=1)
    # config-5 style: tangles with short cycles, wider band
    yield "tangle_cycles", synth.make_case(13, synth.make_graph(13, 30000, chop=32, bubble_every=100, cycle_every=700, tangle_every=5000, tangle_levels=6), 12, 2500, b=20)
    # SimulateReads-style seeds at read position 1 (1-bp backward part: all-N slice, massive ties)
    yield "seed_pos1", synth.make_case(14, synth.make_graph(14, 20000, chop=32, bubble_every=150), 12, 1000, b=10, seed_offsets=(1,))
    # GFA semantics with a 3-bp edge overlap
    c = synth.make_case(15, synth.make_graph(15, 20000, chop=24, snp_every=300), 8, 800, b=10, seed_offsets=(0,))
    c.gfa_overlap = 0
    yield "gfa", c
    # short reads / ragged lengths around the 64-row slice boundary, tiny nodes, error-free and noisy
    yield "ragged_short", synth.make_case(16, synth.make_graph(16, 6000, chop=8, bubble_every=40), 24, 100, b=5, len_jitter=40, seed_offsets=(0, 50))
    yield "wide_band", synth.make_case(17, synth.make_graph(17, 20000, chop=16, bubble_every=60, indel_frac=0.5), 8, 1500, b=100)
    # cyclic band components whose result depends on HOW the reference iterates them: a work list that stops with columns
    # confirmed to 47 of 64 rows (their nodes then carry minScore INT_MAX into the next band selection), and per-node minima
    # left by the last calculateNode call
    yield "cyclic_partial_confirm", fuzz_case(11205)
    yield "cyclic_last_call_min", fuzz_case(1605)
    yield "ramp", synth.make_case(18, synth.make_graph(18, 20000, chop=32, bubble_every=100), 8, 2000, b=5, B=30, errors=(0.08, 0.08, 0.08))
    # -B where the ramp fires: 8 of the 12 reads lose the narrow band somewhere, go back and redo a stretch with the wide one
    # (GraphAligner.h:2648-2719), and the reference still finishes (it does not on every such input, profiles/r02_ramp_fuzz.txt)
    from graphaligner_b200.tools import fuzz
    fuzz.RAMP = True
    yield "ramp_redo", fuzz.make_case(8)[0]
    # -B where the redo leaves a STALE sqrt checkpoint behind (the reference does not rewind its pending checkpoint, GraphAligner.h:2667,
    # 2772-2786) and the reference survives it: the stretch behind that checkpoint is re-computed from a slice of the abandoned
    # narrow-band pass when the trace is taken (getSlicesFromTable, GraphAligner.h:2858-2943), so score and trace are not the ones of
    # the redone forward pass (profiles/r02_ramp_fuzz.txt: read_6 of the first, read_6 and read_10 of the second)
    yield "ramp_stale", fuzz.make_case(30078)[0]
    yield "ramp_stale_long", fuzz.make_case(30025)[0]
    fuzz.RAMP = False


def main():
    only = sys.argv[1:]
    for name, case in cases():
        if only and name not in only:
            continue
        path = os.path.join(OUT, name + ".gacase")
        gacase.write_case(case, path)
        res = subprocess.run([REF, path, "--quiet"], capture_output=True, text=True, check=True)
        lines = [l for l in res.stdout.split("\n") if l and not l.startswith("TIME")]
        with open(os.path.join(OUT, name + ".expected"), "w") as f:
            f.write("\n".join(lines) + "\n")
        ok = sum(1 for l in lines if l.startswith("READ") and "failed=0" in l)
        print("%-20s reads=%d aligned=%d" % (name, len(case.reads), ok))


if __name__ == "__main__":
    main()
