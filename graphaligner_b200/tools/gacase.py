"""Reader/writer for the '.gacase' text test-case format shared by oracle/ref_driver.cpp,
oracle/ga_oracle.cpp, the tests and bench.py (format documented in oracle/ref_driver.cpp)."""


class Case:
    def __init__(self, nodes=None, edges=None, reads=None, b=10, B=0, gfa_overlap=None):
        self.nodes = nodes or []     # [(id, sequence)]
        self.edges = edges or []     # [(from, from_start, to, to_end)]
        self.reads = reads or []     # [(name, sequence, [(node, pos, reverse)])]
        self.b = b
        self.B = B
        self.gfa_overlap = gfa_overlap  # None => vg semantics


def write_case(case, path):
    with open(path, "w") as f:
        if case.gfa_overlap is None:
            f.write("G vg\n")
        else:
            f.write("G gfa %d\n" % case.gfa_overlap)
        for nid, seq in case.nodes:
            f.write("N %d %s\n" % (nid, seq))
        for a, fs, b, te in case.edges:
            f.write("E %d %d %d %d\n" % (a, int(fs), b, int(te)))
        f.write("P %d %d\n" % (case.b, case.B))
        for name, seq, seeds in case.reads:
            f.write("R %s %s %d\n" % (name, seq, len(seeds)))
            for node, pos, rev in seeds:
                f.write("S %d %d %d\n" % (node, pos, int(rev)))


def read_case(path):
    c = Case()
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        p = lines[i].split()
        i += 1
        if not p:
            continue
        if p[0] == "G":
            c.gfa_overlap = int(p[2]) if p[1] == "gfa" else None
        elif p[0] == "N":
            c.nodes.append((int(p[1]), p[2]))
        elif p[0] == "E":
            c.edges.append((int(p[1]), bool(int(p[2])), int(p[3]), bool(int(p[4]))))
        elif p[0] == "P":
            c.b, c.B = int(p[1]), int(p[2])
        elif p[0] == "R":
            n = int(p[3])
            seeds = []
            for _ in range(n):
                q = lines[i].split()
                i += 1
                seeds.append((int(q[1]), int(q[2]), bool(int(q[3]))))
            c.reads.append((p[1], p[2], seeds))
    return c


def parse_ref_output(text):
    """Parse the stdout of oracle/_ref/ref_align (and ga_oracle, same format) into dicts."""
    reads = []
    timing = None
    cur = None
    for line in text.split("\n"):
        if line.startswith("READ "):
            p = line.split()
            cur = {"name": p[1], "mappings": [], "trace": []}
            for kv in p[2:]:
                k, v = kv.split("=")
                cur[k] = int(v, 16) if k in ("th", "mh") else int(v)
            reads.append(cur)
        elif line.startswith("M "):
            cur["mappings"].append(tuple(int(x) for x in line.split()[1:]))
        elif line.startswith("T "):
            cur["trace"].append(tuple(int(x) for x in line.split()[1:]))
        elif line.startswith("TIME "):
            timing = {}
            for kv in line.split()[1:]:
                k, v = kv.split("=")
                timing[k] = float(v)
    return reads, timing


RESULT_KEYS = ("failed", "score", "start", "end", "qpos", "nmap", "ntrace", "th")


def same_result(mine, expected):
    """One read's result dict (Results.as_dicts) against the reference's (parse_ref_output): score, range, query position,
    every mapping and the fingerprint of every trace item."""
    if any(mine[k] != expected[k] for k in RESULT_KEYS):
        return False
    return [tuple(x) for x in mine["mappings"]] == [tuple(x) for x in expected["mappings"]]


def mapping_checksums(read_records, mappings):
    """Per read: sum over its mappings m of (m + 1) * mix(m) mod 2^64 - the order-sensitive checksum ref_align --summary
    prints as mh= (read_records / mappings: the numpy views of api.Results)."""
    import numpy as np
    n = len(read_records)
    out = np.zeros(n, dtype=np.uint64)
    if len(mappings) == 0:
        return out
    with np.errstate(over="ignore"):
        mix = (mappings["node_id"].astype(np.int64).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
               + mappings["is_reverse"].astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)
               + mappings["offset"].astype(np.uint64) * np.uint64(0x165667B19E3779F9)
               + mappings["from_length"].astype(np.int64).astype(np.uint64) * np.uint64(0x27D4EB2F165667C5)
               + mappings["to_length"].astype(np.int64).astype(np.uint64) * np.uint64(0x85EBCA77C2B2AE63))
        off = read_records["mapping_offset"].astype(np.int64)
        cnt = np.where(read_records["failed"] == 0, read_records["n_mappings"], 0).astype(np.int64)
        rank = (np.arange(len(mappings), dtype=np.int64) - np.repeat(off, cnt) + 1).astype(np.uint64) if int(cnt.sum()) == len(mappings) else None
        if rank is None:
            # mappings of failed reads interleaved (never produced by the library, kept for safety): per-read loop
            for i in range(n):
                m = mix[off[i]:off[i] + cnt[i]]
                out[i] = (m * np.arange(1, len(m) + 1, dtype=np.uint64)).sum(dtype=np.uint64)
            return out
        w = mix * rank
        csum = np.concatenate(([np.uint64(0)], np.cumsum(w, dtype=np.uint64)))
        out = csum[off + cnt] - csum[off]
    return out
