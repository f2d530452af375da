"""ctypes binding of include/graphaligner_b200.h.  The shared library is the product; this module only
marshals numpy buffers into the C structs.  There is no fallback: importing works without a GPU (so the
symbol table can be checked), creating an Aligner without one raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GA_LIB: another build of the same library (A/B measurements of kernel variants); there is still no fallback
LIB_PATH = os.environ.get("GA_LIB") or os.path.join(_HERE, "libgraphaligner_b200.so")


class GaBatch(C.Structure):
    _fields_ = [("n_reads", C.c_size_t), ("sequences", C.c_void_p), ("seq_offsets", C.c_void_p), ("names", C.c_void_p),
                ("name_offsets", C.c_void_p), ("seed_offsets", C.c_void_p), ("seed_node", C.c_void_p), ("seed_pos", C.c_void_p),
                ("seed_reverse", C.c_void_p), ("initial_bandwidth", C.c_int32), ("ramp_bandwidth", C.c_int32)]


READ_RESULT = np.dtype([("failed", "<i4"), ("score", "<i4"), ("alignment_start", "<u8"), ("alignment_end", "<u8"),
                        ("query_position", "<i4"), ("flags", "<u4"), ("mapping_offset", "<u8"), ("n_mappings", "<u8"),
                        ("reserved", "<u8"), ("n_trace", "<u8"), ("word_columns", "<u8")])
MAPPING = np.dtype([("node_id", "<i8"), ("offset", "<u4"), ("rank", "<u4"), ("from_length", "<i4"), ("to_length", "<i4"), ("read_start", "<u4"), ("is_reverse", "<u4")])
TRACE_ITEM = np.dtype([("node_id", "<i4"), ("offset", "<u4"), ("readpos", "<u8"), ("reverse", "u1"), ("type", "u1"),
                       ("graph_char", "S1"), ("read_char", "S1"), ("reserved", "<u4")])


class GaStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("streams", "word_columns", "retries", "h2d_bytes", "d2h_bytes", "launches", "graph_bytes", "peq_us", "forward_us", "trace_us")]


EXPORTS = ["ga_create", "ga_destroy", "ga_last_error", "ga_global_error", "ga_graph_new", "ga_graph_add_node", "ga_graph_add_edge",
           "ga_graph_set_dbg_overlap", "ga_graph_finalize", "ga_graph_from_bigraph", "ga_graph_load_vg", "ga_graph_load_gfa",
           "ga_graph_free", "ga_graph_node_count", "ga_graph_size_bp", "ga_graph_edge_count", "ga_graph_upload", "ga_align_batch",
           "ga_stage_batch", "ga_run_staged", "ga_sync", "ga_finish_staged", "ga_staged_free", "ga_cuda_stream", "ga_results_count",
           "ga_results_reads", "ga_results_mappings", "ga_results_read_trace", "ga_results_free", "ga_results_trace_hash", "ga_get_stats",
           "ga_reset_stats", "ga_measure_int32_peak", "ga_pipeline_create", "ga_pipeline_destroy", "ga_pipeline_last_error", "ga_pipeline_depth",
           "ga_pipeline_context", "ga_pipeline_graph_upload", "ga_pipeline_submit", "ga_pipeline_next", "ga_pipeline_in_flight",
           "ga_pipeline_get_stats", "ga_pipeline_reset_stats"]

_lib = None


def load_library():
    """Load libgraphaligner_b200.so; raises if it has not been built (python -m graphaligner_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libgraphaligner_b200.so is missing: run `python graphaligner_b200/build.py` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int32, C.c_uint64
    sig = {
        "ga_create": (vp, [C.c_int]), "ga_destroy": (None, [vp]), "ga_last_error": (C.c_char_p, [vp]), "ga_global_error": (C.c_char_p, []),
        "ga_graph_new": (vp, []), "ga_graph_add_node": (C.c_int, [vp, i32, C.c_char_p, sz, C.c_int]), "ga_graph_add_edge": (C.c_int, [vp, i32, i32]),
        "ga_graph_set_dbg_overlap": (C.c_int, [vp, i32]), "ga_graph_finalize": (C.c_int, [vp]),
        "ga_graph_from_bigraph": (vp, [sz, vp, vp, vp, sz, vp, vp, vp, vp, i32]), "ga_graph_load_vg": (vp, [C.c_char_p]),
        "ga_graph_load_gfa": (vp, [C.c_char_p]), "ga_graph_free": (None, [vp]), "ga_graph_node_count": (sz, [vp]), "ga_graph_size_bp": (sz, [vp]),
        "ga_graph_edge_count": (sz, [vp]), "ga_graph_upload": (C.c_int, [vp, vp]), "ga_align_batch": (vp, [vp, C.POINTER(GaBatch)]),
        "ga_stage_batch": (vp, [vp, C.POINTER(GaBatch)]), "ga_run_staged": (C.c_int, [vp, vp]), "ga_sync": (C.c_int, [vp]),
        "ga_finish_staged": (vp, [vp, vp]), "ga_staged_free": (None, [vp, vp]), "ga_cuda_stream": (vp, [vp]), "ga_results_count": (sz, [vp]),
        "ga_results_reads": (vp, [vp]), "ga_results_mappings": (vp, [vp]), "ga_results_read_trace": (sz, [vp, sz, vp, sz]), "ga_results_free": (None, [vp]),
        "ga_results_trace_hash": (u64, [vp, sz]), "ga_get_stats": (C.c_int, [vp, C.POINTER(GaStats)]), "ga_reset_stats": (C.c_int, [vp]), "ga_measure_int32_peak": (C.c_double, [vp]),
        "ga_pipeline_create": (vp, [C.c_int, C.c_int]), "ga_pipeline_destroy": (None, [vp]), "ga_pipeline_last_error": (C.c_char_p, [vp]),
        "ga_pipeline_depth": (C.c_int, [vp]), "ga_pipeline_context": (vp, [vp, C.c_int]), "ga_pipeline_graph_upload": (C.c_int, [vp, vp]),
        "ga_pipeline_submit": (C.c_int, [vp, C.POINTER(GaBatch)]), "ga_pipeline_next": (vp, [vp]), "ga_pipeline_in_flight": (C.c_int, [vp]),
        "ga_pipeline_get_stats": (C.c_int, [vp, C.POINTER(GaStats)]), "ga_pipeline_reset_stats": (C.c_int, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Graph:
    """Host graph (the reference's AlignmentGraph) built through the C ABI."""

    def __init__(self, handle):
        self._lib = load_library()
        if not handle:
            raise RuntimeError("graph construction failed: " + self._lib.ga_global_error().decode())
        self.handle = handle

    @classmethod
    def from_bigraph(cls, nodes, edges, gfa_overlap=None):
        lib = load_library()
        ids = np.array([n[0] for n in nodes], dtype=np.int64)
        seqs = "".join(n[1] for n in nodes).encode()
        offs = np.zeros(len(nodes) + 1, dtype=np.uint64)
        np.cumsum([len(n[1]) for n in nodes], out=offs[1:])
        seqbuf = np.frombuffer(seqs, dtype=np.uint8) if seqs else np.zeros(1, dtype=np.uint8)
        fr = np.array([e[0] for e in edges], dtype=np.int64)
        fs = np.array([1 if e[1] else 0 for e in edges], dtype=np.uint8)
        to = np.array([e[2] for e in edges], dtype=np.int64)
        te = np.array([1 if e[3] else 0 for e in edges], dtype=np.uint8)
        h = lib.ga_graph_from_bigraph(len(nodes), _ptr(ids), _ptr(seqbuf), _ptr(offs), len(edges), _ptr(fr), _ptr(fs), _ptr(to), _ptr(te),
                                      -1 if gfa_overlap is None else int(gfa_overlap))
        return cls(h)

    @classmethod
    def from_case(cls, case):
        return cls.from_bigraph(case.nodes, case.edges, case.gfa_overlap)

    @classmethod
    def load(cls, path):
        lib = load_library()
        if path.endswith(".vg"):
            return cls(lib.ga_graph_load_vg(path.encode()))
        if path.endswith(".gfa"):
            return cls(lib.ga_graph_load_gfa(path.encode()))
        raise ValueError("Unknown graph type (%s)" % path)

    def node_count(self):
        return self._lib.ga_graph_node_count(self.handle)

    def size_bp(self):
        return self._lib.ga_graph_size_bp(self.handle)

    def edge_count(self):
        return self._lib.ga_graph_edge_count(self.handle)

    def __del__(self):
        if getattr(self, "handle", None):
            self._lib.ga_graph_free(self.handle)
            self.handle = None


class PackedReads:
    """Reads + seeds marshalled into the flat arrays of ga_batch (kept alive with the struct)."""

    def __init__(self, reads, b, B=0, seq_alloc=None):
        # reads: [(name, sequence, [(node, pos, reverse)])]
        # seq_alloc(nbytes) -> uint8 array: where the read bytes are put, e.g. page-locked host memory (a buffer the CUDA runtime
        # knows as pinned is uploaded in place; anything else is first copied into the context's own pinned staging buffer)
        self.n = len(reads)
        self.total_bp = sum(len(r[1]) for r in reads)
        seq = "".join(r[1] for r in reads).encode()
        if seq and seq_alloc is not None:
            self.seq = seq_alloc(len(seq))
            self.seq[:] = np.frombuffer(seq, dtype=np.uint8)
        else:
            self.seq = np.frombuffer(seq, dtype=np.uint8).copy() if seq else np.zeros(1, dtype=np.uint8)
        self.seq_off = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum([len(r[1]) for r in reads], out=self.seq_off[1:])
        names = "".join(r[0] for r in reads).encode()
        self.names = np.frombuffer(names, dtype=np.uint8).copy() if names else np.zeros(1, dtype=np.uint8)
        self.name_off = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum([len(r[0]) for r in reads], out=self.name_off[1:])
        self.seed_off = np.zeros(self.n + 1, dtype=np.uint64)
        np.cumsum([len(r[2]) for r in reads], out=self.seed_off[1:])
        seeds = [s for r in reads for s in r[2]]
        self.seed_node = np.array([s[0] for s in seeds] or [0], dtype=np.int32)
        self.seed_pos = np.array([s[1] for s in seeds] or [0], dtype=np.uint64)
        self.seed_rev = np.array([1 if s[2] else 0 for s in seeds] or [0], dtype=np.uint8)
        self.struct = GaBatch(self.n, _ptr(self.seq), _ptr(self.seq_off), _ptr(self.names), _ptr(self.name_off), _ptr(self.seed_off),
                              _ptr(self.seed_node), _ptr(self.seed_pos), _ptr(self.seed_rev), int(b), int(B))


class Results:
    def __init__(self, lib, handle, names, keepalive=None):
        self._lib = lib
        self.handle = handle
        n = lib.ga_results_count(handle)
        self.reads = self._view(lib.ga_results_reads(handle), READ_RESULT, n)
        # the reads' mapping ranges lie anywhere in the block (mapping_offset): the view covers up to the last one
        ok = self.reads["failed"] == 0
        nm = int((self.reads["mapping_offset"][ok] + self.reads["n_mappings"][ok]).max()) if n and ok.any() else 0
        self.mappings = self._view(lib.ga_results_mappings(handle), MAPPING, nm)
        self.names = names
        self.keepalive = keepalive   # the ga_batch buffers are referenced by the results (lazy trace items)

    @staticmethod
    def _view(ptr, dtype, n):
        if n == 0 or not ptr:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=n)

    def read_trace(self, i):
        """AlignmentResult::trace of read i (materialised on demand)."""
        n = int(self.reads["n_trace"][i])
        buf = np.zeros(max(n, 1), dtype=TRACE_ITEM)
        got = self._lib.ga_results_read_trace(self.handle, i, buf.ctypes.data_as(C.c_void_p), n)
        return buf[:min(n, got)]

    def trace_hash(self, i):
        return int(self._lib.ga_results_trace_hash(self.handle, i))

    def as_dicts(self, with_trace=False, indices=None):
        """Same shape as gacase.parse_ref_output, for differential tests against the oracle (indices: only these reads)."""
        out = []
        for i in (range(len(self.reads)) if indices is None else indices):
            r = self.reads[i]
            failed = int(r["failed"])
            d = {"name": self.names[i] if self.names else str(i), "failed": failed, "asserted": 0,
                 "score": 0 if failed else int(r["score"]), "start": 0 if failed else int(r["alignment_start"]),
                 "end": 0 if failed else int(r["alignment_end"]), "qpos": 0 if failed else int(r["query_position"]),
                 "nmap": 0 if failed else int(r["n_mappings"]), "ntrace": 0 if failed else int(r["n_trace"]),
                 "th": 0 if failed else self.trace_hash(i), "flags": int(r["flags"]), "mappings": [], "trace": []}
            if not failed:
                m = self.mappings[int(r["mapping_offset"]):int(r["mapping_offset"]) + int(r["n_mappings"])]
                d["mappings"] = [(int(x["node_id"]), int(x["is_reverse"]), int(x["offset"]), int(x["from_length"]), int(x["to_length"])) for x in m]
                if with_trace:
                    t = self.read_trace(i)
                    d["trace"] = [(int(x["node_id"]), int(x["offset"]), int(x["reverse"]), int(x["readpos"]), int(x["type"])) for x in t]
            out.append(d)
        return out

    def free(self):
        if self.handle:
            self.reads = self.mappings = None
            self._lib.ga_results_free(self.handle)
            self.handle = None

    def __del__(self):
        self.free()


class Aligner:
    """One GPU context with a replicated graph: the batched form of the reference's AlignOneWay."""

    def __init__(self, graph, device=0):
        self._lib = load_library()
        self.ctx = self._lib.ga_create(int(device))
        if not self.ctx:
            raise RuntimeError("ga_create failed: " + self._lib.ga_global_error().decode())
        self.graph = graph
        if self._lib.ga_graph_upload(self.ctx, graph.handle) != 0:
            raise RuntimeError("ga_graph_upload failed: " + self.last_error())

    def last_error(self):
        return self._lib.ga_last_error(self.ctx).decode()

    def align(self, reads, b=10, B=0):
        packed = reads if isinstance(reads, PackedReads) else PackedReads(reads, b, B)
        h = self._lib.ga_align_batch(self.ctx, C.byref(packed.struct))
        if not h:
            raise RuntimeError("ga_align_batch failed: " + self.last_error())
        names = None if isinstance(reads, PackedReads) else [r[0] for r in reads]
        return Results(self._lib, h, names, keepalive=(packed, self.graph))   # the results refer to the batch buffers and to the graph

    def stage(self, packed):
        h = self._lib.ga_stage_batch(self.ctx, C.byref(packed.struct))
        if not h:
            raise RuntimeError("ga_stage_batch failed: " + self.last_error())
        return h

    def run(self, staged):
        if self._lib.ga_run_staged(self.ctx, staged) != 0:
            raise RuntimeError("ga_run_staged failed: " + self.last_error())

    def sync(self):
        if self._lib.ga_sync(self.ctx) != 0:
            raise RuntimeError("ga_sync failed: " + self.last_error())

    def finish(self, staged, names=None, keepalive=None):
        h = self._lib.ga_finish_staged(self.ctx, staged)
        if not h:
            raise RuntimeError("ga_finish_staged failed: " + self.last_error())
        return Results(self._lib, h, names, keepalive=(keepalive, self.graph))

    def free_staged(self, staged):
        self._lib.ga_staged_free(self.ctx, staged)

    def cuda_stream(self):
        return self._lib.ga_cuda_stream(self.ctx)

    def stats(self):
        s = GaStats()
        self._lib.ga_get_stats(self.ctx, C.byref(s))
        return {n: int(getattr(s, n)) for n, _ in GaStats._fields_}

    def int32_peak(self):
        return float(self._lib.ga_measure_int32_peak(self.ctx))

    def reset_stats(self):
        self._lib.ga_reset_stats(self.ctx)

    def close(self):
        if getattr(self, "ctx", None):
            self._lib.ga_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()


class Pipeline:
    """A stream of batches through `depth` contexts of one GPU (ga_pipeline_*): while one batch's kernel runs, the host
    stages the next and assembles the previous.  Results come back in submission order."""

    def __init__(self, graph, device=0, depth=2):
        self._lib = load_library()
        self.handle = self._lib.ga_pipeline_create(int(device), int(depth))
        if not self.handle:
            raise RuntimeError("ga_pipeline_create failed: " + self._lib.ga_global_error().decode())
        self.graph = graph
        self.depth = int(depth)
        self._pending = []   # (packed, names) of the batches in flight, oldest first
        if self._lib.ga_pipeline_graph_upload(self.handle, graph.handle) != 0:
            raise RuntimeError("ga_pipeline_graph_upload failed: " + self.last_error())

    def last_error(self):
        return self._lib.ga_pipeline_last_error(self.handle).decode()

    def in_flight(self):
        return int(self._lib.ga_pipeline_in_flight(self.handle))

    def submit(self, reads, b=10, B=0):
        packed = reads if isinstance(reads, PackedReads) else PackedReads(reads, b, B)
        rc = self._lib.ga_pipeline_submit(self.handle, C.byref(packed.struct))
        if rc != 0:
            raise RuntimeError("ga_pipeline_submit failed (%d): %s" % (rc, self.last_error()))
        self._pending.append((packed, None if isinstance(reads, PackedReads) else [r[0] for r in reads]))

    def next(self):
        h = self._lib.ga_pipeline_next(self.handle)
        if not self._pending:
            raise RuntimeError("ga_pipeline_next failed: " + self.last_error())
        packed, names = self._pending.pop(0)
        if not h:
            raise RuntimeError("ga_pipeline_next failed: " + self.last_error())
        return Results(self._lib, h, names, keepalive=(packed, self.graph))

    def align_all(self, batches, b=10, B=0):
        """Generator: aligns an iterable of batches, at most `depth` in flight, yielding their Results in order."""
        for batch in batches:
            if self.in_flight() == self.depth:
                yield self.next()
            self.submit(batch, b, B)
        while self.in_flight():
            yield self.next()

    def stats(self):
        s = GaStats()
        self._lib.ga_pipeline_get_stats(self.handle, C.byref(s))
        return {n: int(getattr(s, n)) for n, _ in GaStats._fields_}

    def reset_stats(self):
        self._lib.ga_pipeline_reset_stats(self.handle)

    def close(self):
        if getattr(self, "handle", None):
            self._lib.ga_pipeline_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()
