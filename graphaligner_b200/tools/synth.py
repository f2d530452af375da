"""Seeded synthetic inputs for the parity tests and bench.py: variation graphs
(chopped backbone + SNP / indel bubbles, inversions, tangles with short cycles), reads simulated
as random walks with the reference simulator's error model (reference SimulateReads.cpp:12-41:
per base delete p_del, else substitute p_sub with a uniform base, then with p_ins/10 insert
U{0..19} random bases), and seed hits in the (bigraph node id, read position, reverse) form the
reference consumes (Aligner.cpp:264-271).  Pure numpy; no reference code is executed here.
"""
import numpy as np

from .gacase import Case

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def revcomp(s):
    return "".join(_COMP[c] for c in reversed(s))


def _randseq(rng, n):
    return _BASES[rng.integers(0, 4, size=n)].tobytes().decode()


class Graph:
    """Bidirected graph with integer node ids (>= 2, PickSeedHits.cpp:24 drops ids <= 1)."""

    def __init__(self):
        self.nodes = []          # [(id, seq)] in file order
        self.edges = []          # [(from, from_start, to, to_end)]
        self.seq = {}
        self._succ = None

    def add_node(self, seq):
        nid = len(self.nodes) + 2
        self.nodes.append((nid, seq))
        self.seq[nid] = seq
        return nid

    def add_edge(self, a, b, from_start=False, to_end=False):
        self.edges.append((a, from_start, b, to_end))

    def succ(self):
        """Directed successors over (id, strand) following BigraphToDigraph.cpp:32-56."""
        if self._succ is None:
            s = {}
            for a, fs, b, te in self.edges:
                from_right = (a, 1) if fs else (a, 0)
                from_left = (a, 0) if fs else (a, 1)
                to_right = (b, 1) if te else (b, 0)
                to_left = (b, 0) if te else (b, 1)
                for u, v in ((from_right, to_right), (to_left, from_left)):
                    lst = s.setdefault(u, [])
                    if v not in lst:
                        lst.append(v)
            self._succ = s
        return self._succ

    def strand_seq(self, nid, strand):
        return self.seq[nid] if strand == 0 else revcomp(self.seq[nid])


def make_graph(seed, backbone_len, chop=32, snp_every=0, bubble_every=0, indel_frac=0.2, inversion_every=0,
               tangle_every=0, tangle_levels=20, tangle_width=4, tangle_node=8, cycle_every=0):
    """Chain of elements along a random backbone.

    snp_every      one SNP bubble (two 1-bp alt nodes) per this many bp (config 2 style)
    bubble_every   one bubble per this many bp: SNP with prob 1-indel_frac, else indel
                   (one branch a 1-10 bp node, the other the direct edge) (config 3 style)
    inversion_every  a node traversable on either strand (exercises from_start/to_end edges)
    tangle_every   tangle_levels x tangle_width parallel tangle_node-bp nodes, fully connected
                   level to level, plus back-edges forming 2-3 node cycles (config 5 style)
    cycle_every    a small 2-node cycle hanging on the backbone
    """
    rng = np.random.default_rng(seed)
    g = Graph()
    exits = []   # list of (id, strand) whose right side connects to whatever comes next
    pos = 0
    next_snp = snp_every if snp_every else None
    next_bub = bubble_every if bubble_every else None
    next_inv = inversion_every if inversion_every else None
    next_tan = tangle_every if tangle_every else None
    next_cyc = cycle_every if cycle_every else None

    def connect(prev_exits, entries):
        for (a, sa) in prev_exits:
            for (b, sb) in entries:
                g.add_edge(a, b, from_start=(sa == 1), to_end=(sb == 1))

    def plain(n):
        nonlocal exits, pos
        nid = g.add_node(_randseq(rng, n))
        connect(exits, [(nid, 0)])
        exits = [(nid, 0)]
        pos += n

    while pos < backbone_len:
        nxt = min(x for x in (next_snp, next_bub, next_inv, next_tan, next_cyc, backbone_len) if x is not None)
        # plain chopped nodes up to the next event
        while pos < nxt:
            plain(min(chop, nxt - pos))
        if pos >= backbone_len:
            break
        if next_snp is not None and pos >= next_snp:
            b1, b2 = rng.choice(4, size=2, replace=False)
            n1 = g.add_node("ACGT"[b1])
            n2 = g.add_node("ACGT"[b2])
            connect(exits, [(n1, 0), (n2, 0)])
            exits = [(n1, 0), (n2, 0)]
            pos += 1
            next_snp += snp_every
            plain(min(chop, max(1, backbone_len - pos)))
        elif next_bub is not None and pos >= next_bub:
            if rng.random() < indel_frac:
                n1 = g.add_node(_randseq(rng, int(rng.integers(1, 11))))
                connect(exits, [(n1, 0)])
                exits = exits + [(n1, 0)]
                pos += 1
            else:
                b1, b2 = rng.choice(4, size=2, replace=False)
                n1 = g.add_node("ACGT"[b1])
                n2 = g.add_node("ACGT"[b2])
                connect(exits, [(n1, 0), (n2, 0)])
                exits = [(n1, 0), (n2, 0)]
                pos += 1
            next_bub += bubble_every
            plain(min(chop, max(1, backbone_len - pos)))
        elif next_inv is not None and pos >= next_inv:
            n1 = g.add_node(_randseq(rng, int(rng.integers(min(8, chop), chop + 1))))
            connect(exits, [(n1, 0), (n1, 1)])
            exits = [(n1, 0), (n1, 1)]
            pos += len(g.seq[n1])
            next_inv += inversion_every
            plain(min(chop, max(1, backbone_len - pos)))
        elif next_tan is not None and pos >= next_tan:
            prev_level = None
            for lvl in range(tangle_levels):
                level = [g.add_node(_randseq(rng, tangle_node)) for _ in range(tangle_width)]
                connect(exits, [(n, 0) for n in level])
                if prev_level is not None and lvl % 3 == 1:
                    # back-edges: 2-node cycles between consecutive levels
                    g.add_edge(level[0], prev_level[0])
                    if lvl % 6 == 1 and len(level) > 1:
                        g.add_edge(level[1], prev_level[-1])
                exits = [(n, 0) for n in level]
                prev_level = level
                pos += tangle_node
            next_tan += tangle_every
            plain(min(chop, max(1, backbone_len - pos)))
        elif next_cyc is not None and pos >= next_cyc:
            n1 = g.add_node(_randseq(rng, int(rng.integers(2, 12))))
            n2 = g.add_node(_randseq(rng, int(rng.integers(1, 6))))
            connect(exits, [(n1, 0)])
            g.add_edge(n1, n2)
            g.add_edge(n2, n1)
            if rng.random() < 0.3:
                g.add_edge(n2, n2)
            exits = [(n1, 0)]
            pos += len(g.seq[n1])
            next_cyc += cycle_every
            plain(min(chop, max(1, backbone_len - pos)))
    g.exits = exits
    return g


def introduce_errors(rng, real, p_sub=0.05, p_ins=0.05, p_del=0.05):
    """Vectorised restatement of the reference error model; returns (read, map) where
    map[i] = index in the read at which true base i starts (for seed placement)."""
    arr = np.frombuffer(real.encode(), dtype=np.uint8)
    n = len(arr)
    deleted = rng.random(n) < p_del
    subst = (~deleted) & (rng.random(n) < p_sub)
    out_base = arr.copy()
    out_base[subst] = _BASES[rng.integers(0, 4, size=int(subst.sum()))]
    ins = rng.random(n) < (p_ins / 10.0)
    ins_len = np.where(ins, rng.integers(0, 20, size=n), 0)
    per = (~deleted).astype(np.int64) + ins_len
    starts = np.concatenate(([0], np.cumsum(per)))
    total = int(starts[-1])
    res = _BASES[rng.integers(0, 4, size=total)].copy()
    keep_idx = starts[:-1][~deleted]
    res[keep_idx] = out_base[~deleted]
    return res.tobytes().decode(), starts[:-1]


def simulate_read(rng, g, length, p_sub=0.05, p_ins=0.05, p_del=0.05, start=None, max_tries=50):
    """Random walk from the start of a random node on a random strand (SimulateReads.cpp:49-99
    starts mid-node; starting at a node start keeps the offset-0 seed exact).  Returns
    (read, true_sequence, walk, map) with walk = [(node id, strand, true offset of node start)]."""
    succ = g.succ()
    for _ in range(max_tries):
        if start is None:
            nid = g.nodes[int(rng.integers(0, len(g.nodes)))][0]
            strand = int(rng.integers(0, 2))
        else:
            nid, strand = start
        walk = []
        parts = []
        total = 0
        cur = (nid, strand)
        ok = True
        while total < length:
            s = g.strand_seq(*cur)
            walk.append((cur[0], cur[1], total))
            parts.append(s)
            total += len(s)
            if total >= length:
                break
            nxt = succ.get(cur)
            if not nxt:
                ok = False
                break
            cur = nxt[int(rng.integers(0, len(nxt)))]
        if not ok:
            if start is not None:
                return None
            continue
        real = "".join(parts)[:length]
        read, mp = introduce_errors(rng, real, p_sub, p_ins, p_del)
        if len(read) < 2:
            continue
        return read, real, walk, mp
    return None


def seeds_for(walk, mp, read_len, offsets):
    """PickSeedHits-style seed list: for each requested true offset, the node covering it."""
    seeds = []
    seen = set()
    starts = [w[2] for w in walk]
    for off in offsets:
        if off < 0 or off >= len(mp):
            continue
        k = int(np.searchsorted(starts, off, side="right")) - 1
        nid, strand, _ = walk[k]
        rp = int(mp[off])
        if rp >= read_len:
            rp = read_len - 1
        key = (nid, rp)
        if key in seen:
            continue
        seen.add(key)
        seeds.append((nid, rp, bool(strand)))
    return seeds


def make_case(seed, graph, n_reads, read_len, b=10, B=0, seed_offsets=(0,), decoys=0, errors=(0.05, 0.05, 0.05),
              len_jitter=0):
    rng = np.random.default_rng(seed + 7919)
    reads = []
    i = 0
    misses = 0
    while len(reads) < n_reads:
        ln = read_len if not len_jitter else int(read_len + rng.integers(-len_jitter, len_jitter + 1))
        r = simulate_read(rng, graph, max(2, ln), errors[0], errors[1], errors[2])
        if r is None:
            # the graph has no walk this long from the nodes tried: shorten instead of looping forever
            misses += 1
            if misses % 8 == 0:
                read_len = max(2, read_len // 2)
                len_jitter = min(len_jitter, read_len // 4)
            continue
        read, real, walk, mp = r
        offs = [o if o >= 0 else len(real) + o for o in seed_offsets]
        seeds = seeds_for(walk, mp, len(read), offs)
        for _ in range(decoys):
            nid = graph.nodes[int(rng.integers(0, len(graph.nodes)))][0]
            seeds.append((nid, int(rng.integers(0, len(read))), bool(rng.integers(0, 2))))
        reads.append(("read_%d" % i, read, seeds))
        i += 1
    return Case(list(graph.nodes), list(graph.edges), reads, b, B)


_PAR = {}


def _par_chunk(args):
    seed, n, first = args
    c = make_case(seed, _PAR["graph"], n, **_PAR["kw"])
    return [("read_%d" % (first + k), s, sd) for k, (_, s, sd) in enumerate(c.reads)]


def make_case_parallel(seed, graph, n_reads, read_len, workers=None, chunk=500, read_range=None, **kw):
    """make_case over `workers` forked processes (the full-size configs hold 10^5 reads; one process simulates ~150 reads/s).
    Chunk k of `chunk` reads is make_case(seed + 1000003 * (k + 1), ...), so the read set depends on (seed, chunk) only,
    not on the worker count.  read_range = (lo, hi), multiples of `chunk`: only that part of the read set (a rank's shard)."""
    import multiprocessing as mp
    import os
    workers = workers or os.cpu_count() or 1
    b, B = kw.pop("b", 10), kw.pop("B", 0)
    lo, hi = read_range if read_range is not None else (0, n_reads)
    assert lo % chunk == 0
    graph.succ()   # built once, inherited by the workers
    _PAR["graph"] = graph
    _PAR["kw"] = dict(kw, read_len=read_len, b=b, B=B)
    jobs = [(seed + 1000003 * (first // chunk + 1), min(chunk, hi - first), first) for first in range(lo, hi, chunk)]
    if workers <= 1 or len(jobs) <= 1:
        parts = [_par_chunk(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_par_chunk, jobs)
    _PAR.clear()
    return Case(list(graph.nodes), list(graph.edges), [r for p_ in parts for r in p_], b, B)


def _graph_segment(args):
    seed, length, kw = args
    g = make_graph(seed, length, **kw)
    return g.nodes, g.edges, g.exits


def make_graph_parallel(seed, backbone_len, workers=None, segment=2_000_000, **kw):
    """make_graph built as a chain of independently generated segments of `segment` bp (segment k = make_graph(seed +
    7919 * k, ...)), joined exit -> entry; the result depends on (seed, segment) only.  For the 100 Mbp configs."""
    import multiprocessing as mp
    import os
    workers = workers or os.cpu_count() or 1
    lens = [min(segment, backbone_len - o) for o in range(0, backbone_len, segment)]
    jobs = [(seed + 7919 * k, ln, kw) for k, ln in enumerate(lens)]
    if workers <= 1 or len(jobs) <= 1:
        segs = [_graph_segment(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            segs = pool.map(_graph_segment, jobs)
    g = Graph()
    exits = []
    for nodes, edges, seg_exits in segs:
        off = len(g.nodes)          # ids are 2.. in file order: shift by the nodes already there
        g.nodes.extend((nid + off, seq) for nid, seq in nodes)
        for (a, sa) in exits:       # every segment starts with a plain node (id 2, forward)
            g.edges.append((a, sa == 1, 2 + off, False))
        g.edges.extend((a + off, fs, b + off, te) for a, fs, b, te in edges)
        exits = [(a + off, sa) for a, sa in seg_exits]
    g.seq = dict(g.nodes)
    g.exits = exits
    return g


# ---- the five BASELINE.json configs, scaled by `scale` for tests (scale=1.0 is full size) -------------

def config2(scale=1.0, seed=1):
    g = make_graph(seed, int(5_000_000 * scale), chop=32, snp_every=1000)
    return g, dict(n_reads=max(1, int(10_000 * scale)), read_len=10_000, b=10)


def config3(scale=1.0, seed=2, parallel=False):
    mk = make_graph_parallel if parallel else make_graph
    g = mk(seed, int(100_000_000 * scale), chop=32, bubble_every=100, indel_frac=0.2, inversion_every=5000)
    return g, dict(n_reads=max(1, int(100_000 * scale)), read_len=10_000, b=10, seed_offsets=(0, 5000, -300), decoys=1)


def config4(scale=1.0, seed=4, parallel=False):
    """configs[3]: human-scale GFA graph (GFA semantics, 0-bp edge overlap), 1M reads of 15 kbp sharded over the GPUs."""
    mk = make_graph_parallel if parallel else make_graph
    g = mk(seed, int(3_000_000_000 * scale), chop=32, snp_every=1000, bubble_every=5000, indel_frac=0.3)
    return g, dict(n_reads=max(1, int(1_000_000 * scale)), read_len=15_000, b=10, gfa_overlap=0)


def config5(scale=1.0, seed=5, parallel=False):
    mk = make_graph_parallel if parallel else make_graph
    g = mk(seed, int(100_000_000 * scale), chop=32, bubble_every=100, indel_frac=0.2, tangle_every=250_000)
    return g, dict(n_reads=max(1, int(2_000 * scale)), read_len=50_000, b=10)
