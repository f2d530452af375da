"""Benchmark-harness tools of the reference, restated over the protobuf-free vg/GAM codec:

    python -m graphaligner_b200.tools.evaltools simulate graph.vg truth.gam reads.fastq N LENGTH SUB INS seeds.gam DEL
    python -m graphaligner_b200.tools.evaltools pickseeds out.gam MAXSEEDS seeds1.gam [seeds2.gam ...]
    python -m graphaligner_b200.tools.evaltools compare truth.gam predicted.gam graph.vg

Argument order and semantics follow SimulateReads.cpp:155-203, PickSeedHits.cpp:17-36 and CompareAlignments.cpp:51-100.
These are host-side tools around the hot path; nothing here touches the GPU."""
import sys

import numpy as np

from . import synth, vgio


def encode_alignment(name, sequence, mappings, score=0, query_position=0):
    """Alignment{sequence, path{mapping{position{node_id, offset, is_reverse}}}, name, score, query_position}"""
    path = bytearray()
    for node_id, offset, is_reverse in mappings:
        pos = bytearray()
        vgio._put_int(pos, 1, node_id)
        vgio._put_int(pos, 2, offset)
        vgio._put_int(pos, 4, 1 if is_reverse else 0)
        mapping = bytearray()
        vgio._put_bytes(mapping, 1, bytes(pos), True)
        vgio._put_bytes(path, 2, bytes(mapping), True)
    out = bytearray()
    vgio._put_bytes(out, 1, sequence.encode())
    vgio._put_bytes(out, 2, bytes(path), True)
    vgio._put_bytes(out, 3, name.encode())
    vgio._put_int(out, 6, score)
    vgio._put_int(out, 7, query_position)
    return bytes(out)


def graph_from_vg(path):
    nodes, edges = vgio.load_vg_graph(path)
    g = synth.Graph()
    g.nodes = [(n["id"], n["sequence"]) for n in nodes]
    g.edges = [(e["from"], e["from_start"], e["to"], e["to_end"]) for e in edges]
    g.seq = dict(g.nodes)
    return g


def simulate_reads(graph, n_reads, length, p_sub, p_ins, p_del, seed=0):
    """SimulateReads.cpp:49-152: random walks with substitution / insertion / deletion errors.  Returns
    (truth records, fastq records, seed records): the truth is the error-free sequence on its node path, the seed is the
    walk's first node.  Unlike the reference (which starts mid-node and writes query_position 1), walks start at a node
    start and the seed sits at read offset 0, where it is exact."""
    rng = np.random.default_rng(seed)
    truth, fastq, seeds = [], [], []
    while len(fastq) < n_reads:
        r = synth.simulate_read(rng, graph, length, p_sub, p_ins, p_del)
        if r is None:
            raise RuntimeError("the graph has no walk of %d bp" % length)
        read, real, walk, _ = r
        name = "read_%d" % int(rng.integers(0, 1 << 31))
        truth.append(encode_alignment(name, real, [(nid, 0, bool(strand)) for nid, strand, _ in walk]))
        fastq.append((name, read))
        seeds.append(vgio.encode_seed(name, walk[0][0], 0, bool(walk[0][1])))
    return truth, fastq, seeds


def pick_seed_hits(seed_files, max_seeds):
    """PickSeedHits.cpp:9-36: per read name, the first max_seeds distinct (node, query_position) hits over the files in order;
    hits on node ids <= 1 are dropped; output grouped by read name in sorted order (std::map)."""
    picked = {}
    for path in seed_files:
        for a in vgio.load_gam(path):
            if not a["path"] or a["path"][0]["position"]["node_id"] <= 1:
                continue
            have = picked.setdefault(a["name"], [])
            key = (a["path"][0]["position"]["node_id"], a["query_position"])
            if any(key == (h["path"][0]["position"]["node_id"], h["query_position"]) for h in have):
                continue
            if len(have) < max_seeds:
                have.append(a)
    out = []
    for name in sorted(picked):
        for a in picked[name]:
            p = a["path"][0]["position"]
            out.append(vgio.encode_seed(name, p["node_id"], a["query_position"], p["is_reverse"]))
    return out


def alignment_identity(real, predicted, node_sizes):
    """CompareAlignments.cpp:13-44: (common, false negative, false positive) bp over the node SETS of the two paths."""
    left = set(m["position"]["node_id"] for m in real["path"])
    right = set(m["position"]["node_id"] for m in predicted["path"])
    common = sum(node_sizes[n] for n in left & right)
    fn = sum(node_sizes[m["position"]["node_id"]] for m in real["path"]) - common
    fp = sum(node_sizes[m["position"]["node_id"]] for m in predicted["path"]) - common
    return common, fn, fp


def compare_alignments(truth_path, predicted_path, graph_path, out=sys.stdout):
    nodes, _ = vgio.load_vg_graph(graph_path)
    node_sizes = {n["id"]: len(n["sequence"]) for n in nodes}
    real = {a["name"]: a for a in vgio.load_gam(truth_path)}
    predicted = {a["name"]: a for a in vgio.load_gam(predicted_path)}
    good = bad = 0
    for name in sorted(real):
        if name not in predicted:
            bad += 1
            continue
        p = predicted[name]
        c, fn, fp = alignment_identity(real[name], p, node_sizes)
        ident = c / (c + fn + fp) if c + fn + fp else 0.0
        n = max(1, len(p["sequence"]))
        out.write("%s: %dbp common, %dbp false negative, %dbp false positive (%g) %d mismatches, read length %d (%g)\n"
                  % (name, c, fn, fp, ident, p["score"], len(p["sequence"]), p["score"] / n))
        if ident < 0.7:
            bad += 1
        else:
            good += 1
    bad += sum(1 for name in predicted if name not in real)
    out.write("good matches: %d\nbad matches: %d\n" % (good, bad))
    return good, bad


def main(argv):
    if len(argv) < 2:
        sys.stderr.write(__doc__)
        return 2
    cmd, a = argv[1], argv[2:]
    if cmd == "simulate" and len(a) >= 9:
        g = graph_from_vg(a[0])
        truth, fastq, seeds = simulate_reads(g, int(a[3]), int(a[4]), float(a[5]), float(a[6]), float(a[8]))
        vgio.write_stream(a[1], truth)
        vgio.write_stream(a[7], seeds)
        with open(a[2], "w") as f:
            for name, seq in fastq:
                f.write("@%s\n%s\n+\n%s\n" % (name, seq, "!" * len(seq)))
        return 0
    if cmd == "pickseeds" and len(a) >= 3:
        vgio.write_stream(a[0], pick_seed_hits(a[2:], int(a[1])))
        return 0
    if cmd == "compare" and len(a) >= 3:
        compare_alignments(a[0], a[1], a[2])
        return 0
    sys.stderr.write(__doc__)
    return 2


if __name__ == "__main__":
    sys.exit(main(sys.argv))
