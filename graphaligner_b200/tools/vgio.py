"""Protobuf-free reader/writer for the vg "stream" container and the few vg.proto messages on the
aligner's I/O surface (Graph/Node/Edge, Alignment/Path/Mapping/Position/Edit).

Container (reference stream.hpp:24-51 write, :69-111 read): concatenated gzip members holding
  varint64 count, then count x { varint32 length, message bytes }.
Field numbers (reference vg.pb.h:149-960): Graph{1 node,2 edge}; Node{1 sequence,2 name,3 id};
Edge{1 from,2 to,3 from_start,4 to_end,5 overlap}; Edit{1 from_length,2 to_length,3 sequence};
Mapping{1 position,2 edit,5 rank}; Position{1 node_id,2 offset,4 is_reverse};
Path{1 name,2 mapping}; Alignment{1 sequence,2 path,3 name,4 quality,6 score,7 query_position}.
Used by tests/bench tooling; the C++ host code has its own codec (csrc/vgcodec.cpp).
"""
import gzip
import io


def _varint(buf, pos):
    shift = 0
    val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _fields(buf):
    pos = 0
    out = []
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        out.append((fn, wt, v))
    return out


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def read_stream(path):
    """Yield raw message bytes of every record in a vg stream file."""
    with open(path, "rb") as f:
        raw = gzip.GzipFile(fileobj=io.BytesIO(f.read())).read()
    pos = 0
    while pos < len(raw):
        count, pos = _varint(raw, pos)
        for _ in range(count):
            ln, pos = _varint(raw, pos)
            yield raw[pos:pos + ln]
            pos += ln


def parse_graph(msg):
    nodes, edges = [], []
    for fn, _, v in _fields(msg):
        if fn == 1:
            n = {"sequence": "", "name": "", "id": 0}
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    n["sequence"] = v2.decode()
                elif f2 == 2:
                    n["name"] = v2.decode()
                elif f2 == 3:
                    n["id"] = _signed(v2)
            nodes.append(n)
        elif fn == 2:
            e = {"from": 0, "to": 0, "from_start": False, "to_end": False, "overlap": 0}
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    e["from"] = _signed(v2)
                elif f2 == 2:
                    e["to"] = _signed(v2)
                elif f2 == 3:
                    e["from_start"] = bool(v2)
                elif f2 == 4:
                    e["to_end"] = bool(v2)
                elif f2 == 5:
                    e["overlap"] = _signed(v2)
            edges.append(e)
    return nodes, edges


def load_vg_graph(path):
    nodes, edges = [], []
    for msg in read_stream(path):
        n, e = parse_graph(msg)
        nodes += n
        edges += e
    return nodes, edges


def parse_position(msg):
    p = {"node_id": 0, "offset": 0, "is_reverse": False}
    for fn, _, v in _fields(msg):
        if fn == 1:
            p["node_id"] = _signed(v)
        elif fn == 2:
            p["offset"] = _signed(v)
        elif fn == 4:
            p["is_reverse"] = bool(v)
    return p


def parse_mapping(msg):
    m = {"position": {"node_id": 0, "offset": 0, "is_reverse": False}, "edits": [], "rank": 0}
    for fn, _, v in _fields(msg):
        if fn == 1:
            m["position"] = parse_position(v)
        elif fn == 2:
            e = {"from_length": 0, "to_length": 0, "sequence": ""}
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    e["from_length"] = _signed(v2)
                elif f2 == 2:
                    e["to_length"] = _signed(v2)
                elif f2 == 3:
                    e["sequence"] = v2.decode()
            m["edits"].append(e)
        elif fn == 5:
            m["rank"] = _signed(v)
    return m


def parse_alignment(msg):
    a = {"sequence": "", "name": "", "quality": b"", "score": 0, "query_position": 0, "path": []}
    for fn, _, v in _fields(msg):
        if fn == 1:
            a["sequence"] = v.decode()
        elif fn == 2:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    a["path"].append(parse_mapping(v2))
        elif fn == 3:
            a["name"] = v.decode()
        elif fn == 4:
            a["quality"] = v
        elif fn == 6:
            a["score"] = _signed(v) if v < (1 << 63) else _signed(v)
        elif fn == 7:
            a["query_position"] = _signed(v)
    return a


def load_gam(path):
    return [parse_alignment(m) for m in read_stream(path)]


def load_fastq(path):
    """Same record rules as reference fastqloader.cpp:6-29 (4-line records, '@' header)."""
    out = []
    with open(path) as f:
        lines = [l.rstrip("\r\n") for l in f]
    i = 0
    while i < len(lines):
        if not lines[i].startswith("@"):
            i += 1
            continue
        out.append((lines[i][1:], lines[i + 1]))
        i += 4
    return out


# ---- writers (tests build .vg / .gam inputs for the loaders and the CLI with these) ----------------------------

def _put_varint(out, v):
    if v < 0:
        v += 1 << 64
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _put_int(out, field, v):
    if v:
        _put_varint(out, field << 3)
        _put_varint(out, int(v))


def _put_bytes(out, field, b, always=False):
    if b or always:
        _put_varint(out, field << 3 | 2)
        _put_varint(out, len(b))
        out.extend(b)


def encode_graph(nodes, edges):
    """nodes: [(id, sequence)], edges: [(from, from_start, to, to_end)]"""
    out = bytearray()
    for nid, seq in nodes:
        m = bytearray()
        _put_bytes(m, 1, seq.encode())
        _put_int(m, 3, nid)
        _put_bytes(out, 1, bytes(m), True)
    for a, fs, b, te in edges:
        m = bytearray()
        _put_int(m, 1, a)
        _put_int(m, 2, b)
        _put_int(m, 3, 1 if fs else 0)
        _put_int(m, 4, 1 if te else 0)
        _put_bytes(out, 2, bytes(m), True)
    return bytes(out)


def encode_seed(name, node_id, query_position, is_reverse):
    """A seed hit as the aligner reads it: Alignment{name, query_position, path.mapping[0].position}"""
    pos = bytearray()
    _put_int(pos, 1, node_id)
    _put_int(pos, 4, 1 if is_reverse else 0)
    mapping = bytearray()
    _put_bytes(mapping, 1, bytes(pos), True)
    path = bytearray()
    _put_bytes(path, 2, bytes(mapping), True)
    out = bytearray()
    _put_bytes(out, 2, bytes(path), True)
    _put_bytes(out, 3, name.encode())
    _put_int(out, 7, query_position)
    return bytes(out)


def write_stream(path, records, group=1000):
    """gzip members of: varint count, then (varint length, message) x count (stream.hpp:24-51)"""
    with open(path, "wb") as f:
        for i in range(0, len(records), group):
            chunk = records[i:i + group]
            raw = bytearray()
            _put_varint(raw, len(chunk))
            for r in chunk:
                _put_varint(raw, len(r))
                raw.extend(r)
            f.write(gzip.compress(bytes(raw)))
        if not records:
            f.write(gzip.compress(b""))


def write_case_files(case, prefix):
    """Writes <prefix>.vg (or .gfa), <prefix>.fastq, <prefix>_seeds.gam for the reference-style command line."""
    if case.gfa_overlap is None:
        graph_path = prefix + ".vg"
        write_stream(graph_path, [encode_graph(case.nodes, case.edges)])
    else:
        graph_path = prefix + ".gfa"
        with open(graph_path, "w") as f:
            for nid, seq in case.nodes:
                f.write("S\t%d\t%s\n" % (nid, seq))
            for a, fs, b, te in case.edges:
                f.write("L\t%d\t%s\t%d\t%s\t%dM\n" % (a, "-" if fs else "+", b, "-" if te else "+", case.gfa_overlap))
    with open(prefix + ".fastq", "w") as f:
        for name, seq, _ in case.reads:
            f.write("@%s\n%s\n+\n%s\n" % (name, seq, "!" * len(seq)))
    seeds = [encode_seed(name, node, pos, rev) for name, _, ss in case.reads for node, pos, rev in ss]
    write_stream(prefix + "_seeds.gam", seeds)
    return graph_path, prefix + ".fastq", prefix + "_seeds.gam"
