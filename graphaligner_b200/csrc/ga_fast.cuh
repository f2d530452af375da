// Forward pass, small-band kernel body: the same slices as ga_core.cuh's general path, for the streams whose bands are
// acyclic, hold at most GAF_NODES nodes / GAF_COLS columns and run with a fixed bandwidth (no -B ramp).  That is every stream
// of a chopped variation graph at moderate bandwidths; a stream that leaves these limits reports an overflow status and is
// run again by the general kernel.
//
// What is different from the general path is where the state lives and how the lanes of a warp stay together:
//   * everything a slice reads back - the band's node table with each node's graph record (length, first two in- and
//     out-neighbours, sequence chunk), the frozen end state of the previous slice's columns (16-bit), the priority queue,
//     the match words - is in shared memory, addressed as shared memory; global memory only receives the history
//     (columns, node lists, slice headers) and supplies one 32-byte node record per node that ENTERS a band;
//   * a slice's columns are one flat loop: every lane advances one column per iteration whatever node it is in, a node
//     that continues its only in-neighbour (a chain) starts inside the same loop body from the registers of the previous
//     iteration, and only the other node starts (sources, bubbles closing) take a branch of their own;
//   * node lookups are linear scans of the (at most 16-entry) shared-memory node tables; there are no hash tables.
// The arithmetic (word step, min-merge, band selection order incl. the std::unordered_map / std::priority_queue order
// emulation, HMM, stop rule) is ga_core.cuh's; results are bit-identical.
#ifndef GA_FAST_CUH
#define GA_FAST_CUH
#include "ga_core.cuh"

#define GAF_NODES 16u
#define GAF_COLS 256u
#define GAF_HEAP 32u
#define GAF_NONE 0xffu
#define GAF_NOPCS 0xfffu
#ifndef GAF_STEPS
#define GAF_STEPS 2         /* columns of a node a lane takes per pass of the column loop (measured: 1 -> 8.10 ms, 2 -> 7.55 ms, 3 -> 7.77 ms, 4 -> 7.61 ms) */
#endif
#define GAF_ORD_CHAIN 0x100u
#define GAF_ORD_NOUP 0x200u

// One warp's shared memory: [..][lane] so that the lanes of a warp hit different banks
template <int S>
struct GaFastShared
{
	uint64_t heap[GAF_HEAP][S];        // band selection: priority queue; fill: free
	uint64_t eq[2][4][S];              // match words of this slice and (being fetched) the next one
	uint32_t aux[2][S];                // per slice: exact code of the read character above the slice | IUPAC mask of the read's first character << 4
	// node tables of the current and the previous slice
	uint32_t nodeId[2][GAF_NODES][S];
	uint32_t lenDeg[2][GAF_NODES][S];  // ga_node_rec::lenDeg
	uint32_t csPcs[2][GAF_NODES][S];   // first column | first column in the previous slice << 16 (GAF_NOPCS: not in the previous band)
	int32_t nodeMin[2][GAF_NODES][S];
	uint32_t chunk[2][GAF_NODES][S];   // ga_node_rec::seqChunk
	uint32_t nbr[2][4][GAF_NODES][S];  // in[0], in[1], out[0], out[1]
	uint32_t scratch[64][S];           // selection: unordered_map emulation (32 buckets, 16 links, 16 order); fill: in-slots, order
	uint16_t tiny[2][GAF_COLS][S];     // frozen end state per column, current / previous slice (ga_tiny_ld<.., true>)
};

// a lane's view of its warp's block: sh.X[..][lane].  (No table of pointers: indexing one by the slice parity would push
// the pointers into local memory and turn every access into a generic load.)
template <int S>
struct GaFastLane
{
	GaFastShared<S>& sh;
	uint32_t lane;
	GA_DEV GaFastLane(GaFastShared<S>& sh_, uint32_t lane_) : sh(sh_), lane(lane_) {}
	GA_DEV uint64_t* heap() const { return &sh.heap[0][lane]; }
	GA_DEV uint32_t* scratch() const { return &sh.scratch[0][lane]; }
	GA_DEV uint32_t* ids(int t) const { return &sh.nodeId[t][0][lane]; }
	GA_DEV uint32_t tinyLd(int t, uint32_t idx, int32_t ref) const
	{
		const uint32_t v = sh.tiny[t][idx][lane];
		const uint32_t score = (uint32_t)ref + (((v >> 3) - (uint32_t)ref) & 0x1fffu);
		return (score << 3) | (v & 7u);
	}
};

// asynchronous 8-byte / 4-byte copy global -> shared (the next slice's match words arrive while this slice is computed)
#if defined(__CUDACC__)
GA_DEV void ga_cp_async8(void* dstShared, const void* src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(dstShared)), "l"(src));
}
GA_DEV void ga_cp_async4(void* dstShared, const void* src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(dstShared)), "l"(src));
}
GA_DEV void ga_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#else
GA_DEV void ga_cp_async8(void* dst, const void* src) { memcpy(dst, src, 8); }
GA_DEV void ga_cp_async4(void* dst, const void* src) { memcpy(dst, src, 4); }
GA_DEV void ga_cp_async_wait() {}
#endif

// node -> slot in a node table, or GAF_NONE.  No early exit: the lanes of a warp scan tables of about the same size and
// stay together (a loop that each lane leaves at its own hit runs at one or two lanes per instruction).
template <int S>
GA_DEV uint32_t gaf_find(const uint32_t* ids, uint32_t n, uint32_t node)
{
	uint32_t r = GAF_NONE;
	for (uint32_t i = 0; i < n; i++) r = ids[(size_t)i * S] == node ? i : r;
	return r;
}

// key % buckets for the bucket counts a table of up to GAF_NODES keys goes through (13, 29): constant divisors
GA_DEV uint32_t gaf_mod(uint32_t key, uint32_t nb)
{
	return nb == 13u ? key % 13u : (nb == 29u ? key % 29u : (nb == 1u ? 0u : key % nb));
}

// ga_umap_order for at most GAF_NODES keys in shared memory (iteration order of a libstdc++ unordered_map filled key by
// key, see ga_core.cuh): ubkt = scratch[0..32), unext = scratch[32..48), uorder = scratch[48..64)
template <int S>
GA_DEV void gaf_umap_order(const GaUmapSchedule& sch, const uint32_t* keys, uint32_t* scratch, uint32_t n)
{
	uint32_t* ubkt = scratch;
	uint32_t* unext = scratch + (size_t)32 * S;
	uint32_t* uorder = scratch + (size_t)48 * S;
	uint32_t bktCount = 1;
	uint32_t head = GA_UNIL;
	uint32_t si = 0;
	ubkt[0] = GA_UB_EMPTY;
	for (uint32_t i = 0; i < n; i++)
	{
		if (si < sch.n && i + 1 == sch.threshold[si])
		{
			uint32_t nb = sch.buckets[si++];
			for (uint32_t b = 0; b < nb; b++) ubkt[(size_t)b * S] = GA_UB_EMPTY;
			uint32_t p = head;
			head = GA_UNIL;
			uint32_t bbeginBkt = 0;
			while (p != GA_UNIL)
			{
				uint32_t nxt = unext[(size_t)p * S];
				uint32_t b = gaf_mod(keys[(size_t)p * S], nb);
				uint32_t before = ubkt[(size_t)b * S];
				if (before == GA_UB_EMPTY)
				{
					unext[(size_t)p * S] = head;
					bool hadNext = head != GA_UNIL;
					head = p;
					ubkt[(size_t)b * S] = GA_UB_BEGIN;
					if (hadNext) ubkt[(size_t)bbeginBkt * S] = p;
					bbeginBkt = b;
				}
				else
				{
					uint32_t after = before == GA_UB_BEGIN ? head : unext[(size_t)before * S];
					unext[(size_t)p * S] = after;
					if (before == GA_UB_BEGIN) head = p; else unext[(size_t)before * S] = p;
				}
				p = nxt;
			}
			bktCount = nb;
		}
		uint32_t key = keys[(size_t)i * S];
		uint32_t b = gaf_mod(key, bktCount);
		uint32_t before = ubkt[(size_t)b * S];
		if (before != GA_UB_EMPTY)
		{
			uint32_t after = before == GA_UB_BEGIN ? head : unext[(size_t)before * S];
			unext[(size_t)i * S] = after;
			if (before == GA_UB_BEGIN) head = i; else unext[(size_t)before * S] = i;
		}
		else
		{
			unext[(size_t)i * S] = head;
			if (head != GA_UNIL) ubkt[(size_t)gaf_mod(keys[(size_t)head * S], bktCount) * S] = i;
			head = i;
			ubkt[(size_t)b * S] = GA_UB_BEGIN;
		}
	}
	uint32_t k = 0;
	for (uint32_t p = head; p != GA_UNIL; p = unext[(size_t)p * S]) uorder[(size_t)(k++) * S] = p;
}

struct GaFastRec
{
	uint32_t seqChunk, lenDeg, inOff, outOff, in0, in1, out0, out1;
};

GA_DEV GaFastRec gaf_load_rec(const ga_graph_view& g, uint32_t node)
{
	const uint4* p = (const uint4*)(g.nodeRec + node);
	const uint4 a = p[0], b = p[1];
	GaFastRec r;
	r.seqChunk = a.x; r.lenDeg = a.y; r.inOff = a.z; r.outOff = a.w;
	r.in0 = b.x; r.in1 = b.y; r.out0 = b.z; r.out1 = b.w;
	return r;
}

// appends a node to the current band's table; false = the band left the small-band limits
template <int S>
GA_DEV bool gaf_band_add(const GaFastLane<S>& fl, int tc, GaStreamState& st, uint32_t& nc, uint32_t& ncols, uint32_t node, uint32_t lenDeg, uint32_t pcs, uint32_t chunk, uint32_t in0, uint32_t in1, uint32_t out0, uint32_t out1)
{
	const uint32_t len = GA_REC_LEN(lenDeg);
	if (nc >= GAF_NODES) { st.status = GA_ERR_NODE_OVERFLOW; return false; }
	if (ncols + len > GAF_COLS) { st.status = GA_ERR_COL_OVERFLOW; return false; }
	fl.sh.nodeId[tc][nc][fl.lane] = node;
	fl.sh.lenDeg[tc][nc][fl.lane] = lenDeg;
	fl.sh.csPcs[tc][nc][fl.lane] = ncols | (pcs << 16);
	fl.sh.chunk[tc][nc][fl.lane] = chunk;
	fl.sh.nbr[tc][0][nc][fl.lane] = in0;
	fl.sh.nbr[tc][1][nc][fl.lane] = in1;
	fl.sh.nbr[tc][2][nc][fl.lane] = out0;
	fl.sh.nbr[tc][3][nc][fl.lane] = out1;
	nc++;
	ncols += len;
	return true;
}

// pushes the out-neighbours of a band node (slot of table t) with the given priority
template <int S>
GA_DEV bool gaf_push_out(const ga_graph_view& g, const GaFastLane<S>& fl, int t, uint32_t slot, GaStreamState& st, uint32_t& heapN, uint32_t prio)
{
	const uint32_t lenDeg = fl.sh.lenDeg[t][slot][fl.lane];
	const uint32_t outDeg = GA_REC_OUTDEG(lenDeg);
	if (outDeg <= 2)
	{
		for (uint32_t e = 0; e < outDeg; e++)
		{
			if (heapN >= GAF_HEAP) { st.status = GA_ERR_QUEUE_OVERFLOW; return false; }
			ga_heap_push<S>(fl.heap(), heapN, ((uint64_t)prio << 32) | fl.sh.nbr[t][2 + e][slot][fl.lane]);
		}
		return true;
	}
	// longer lists: the CSR (the degree field saturates at 15)
	const uint32_t node = fl.sh.nodeId[t][slot][fl.lane];
	for (uint32_t e = g.outOff[node], eEnd = g.outOff[node + 1]; e < eEnd; e++)
	{
		if (heapN >= GAF_HEAP) { st.status = GA_ERR_QUEUE_OVERFLOW; return false; }
		ga_heap_push<S>(fl.heap(), heapN, ((uint64_t)prio << 32) | g.outAdj[e]);
	}
	return true;
}

// Band selection for slice s from slice s-1 (projectForwardFromMinScore, GraphAligner.h:1110-1159), cf. ga_select_band.
// Called by ALL lanes of the warp (run = this lane has a slice to do): the two loops are warp-wide with a vote per round, so
// that the lanes - which each walk their own few nodes and queue entries through data-dependent branches - meet again every
// round instead of finishing the whole selection one lane group after the other.
template <int S>
GA_DEV int gaf_select_band(const ga_graph_view& g, const GaUmapSchedule& sch, const GaFastLane<S>& fl, GaStreamState& st, bool run, int bandwidth, int tp, uint32_t pNodesIn, uint32_t& ncolsOut)
{
	const int tc = tp ^ 1;
	const int32_t expand = bandwidth + 64;
	const uint32_t pNodes = run ? pNodesIn : 0;
	uint32_t nc = 0, ncols = 0, heapN = 0;
	bool ok = true;
	gaf_umap_order<S>(sch, fl.ids(tp), fl.scratch(), pNodes);
	GA_SYNCWARP();
	const uint32_t* uorder = fl.scratch() + (size_t)48 * S;
	const uint32_t rounds = GA_WARP_MAX(pNodes);
	for (uint32_t it = 0; it < rounds; it++)
	{
		if (it < pNodes && ok)
		{
			const uint32_t i = uorder[(size_t)it * S];
			const int32_t nodeMin = fl.sh.nodeMin[tp][i][fl.lane];
			if (nodeMin <= st.prevMin + bandwidth)
			{
				const uint32_t lenDeg = fl.sh.lenDeg[tp][i][fl.lane];
				const uint32_t len = GA_REC_LEN(lenDeg);
				const uint32_t pcs = fl.sh.csPcs[tp][i][fl.lane] & 0xffffu;
				ok = gaf_band_add<S>(fl, tc, st, nc, ncols, fl.sh.nodeId[tp][i][fl.lane], lenDeg, pcs, fl.sh.chunk[tp][i][fl.lane], fl.sh.nbr[tp][0][i][fl.lane], fl.sh.nbr[tp][1][i][fl.lane],
					fl.sh.nbr[tp][2][i][fl.lane], fl.sh.nbr[tp][3][i][fl.lane]);
				if (ok)
				{
					const int32_t endscore = ga_tiny_score(fl.tinyLd(tp, pcs + len - 1, st.prevMin));
					if (endscore <= st.prevMin + expand) ok = gaf_push_out<S>(g, fl, tp, i, st, heapN, (uint32_t)(endscore - st.prevMin + 1));
				}
			}
		}
		GA_SYNCWARP();
	}
	while (true)
	{
		bool more = ok && heapN > 0;
		uint64_t top = 0;
		if (more)
		{
			top = fl.heap()[0];
			more = (int32_t)(top >> 32) <= expand;
		}
		if (!GA_WARP_ANY(more)) break;
		if (more)
		{
			const int32_t prio = (int32_t)(top >> 32);
			ga_heap_pop<S>(fl.heap(), heapN);
			const uint32_t node = (uint32_t)top;
			if (gaf_find<S>(fl.ids(tc), nc, node) == GAF_NONE)
			{
				// not kept, but it may still sit in the previous band (its minimum was outside the bandwidth)
				const uint32_t pslot = gaf_find<S>(fl.ids(tp), pNodes, node);
				const uint32_t slot = nc;
				if (pslot != GAF_NONE)
				{
					ok = gaf_band_add<S>(fl, tc, st, nc, ncols, node, fl.sh.lenDeg[tp][pslot][fl.lane], fl.sh.csPcs[tp][pslot][fl.lane] & 0xffffu, fl.sh.chunk[tp][pslot][fl.lane],
						fl.sh.nbr[tp][0][pslot][fl.lane], fl.sh.nbr[tp][1][pslot][fl.lane], fl.sh.nbr[tp][2][pslot][fl.lane], fl.sh.nbr[tp][3][pslot][fl.lane]);
				}
				else
				{
					const GaFastRec r = gaf_load_rec(g, node);
					ok = gaf_band_add<S>(fl, tc, st, nc, ncols, node, r.lenDeg, GAF_NOPCS, r.seqChunk, r.in0, r.in1, r.out0, r.out1);
				}
				if (ok)
				{
					const uint32_t len = GA_REC_LEN(fl.sh.lenDeg[tc][slot][fl.lane]);
					ok = gaf_push_out<S>(g, fl, tc, slot, st, heapN, (uint32_t)(prio + (int32_t)len));
				}
			}
		}
	}
	ncolsOut = ncols;
	return ok ? (int)nc : -1;
}

// In-neighbours of every band node that are in the current or the previous band (at most two per node in this kernel), and
// a topological order of the band.  scratch[0..16) = per node: cur slot 0 | prev slot 0 << 8 | cur slot 1 << 16 | prev slot 1
// << 24 (GAF_NONE = absent; pair 0 is filled first); scratch[16..32) = evaluation order; scratch[32..48) = pending counts.
// Returns false when the stream has to leave the small-band kernel (third in-neighbour, cycle).
template <int S>
GA_DEV bool gaf_resolve(const ga_graph_view& g, const GaFastLane<S>& fl, GaStreamState& st, int tc, uint32_t nc, uint32_t pNodes)
{
	const int tp = tc ^ 1;
	uint32_t* inSlots = fl.scratch();
	uint32_t* order = fl.scratch() + (size_t)16 * S;
	uint32_t* pend = fl.scratch() + (size_t)32 * S;
	uint32_t ready = 0;
	bool ok = true;
	for (uint32_t i = 0; i < nc; i++)
	{
		const uint32_t inDeg = GA_REC_INDEG(fl.sh.lenDeg[tc][i][fl.lane]);
		uint32_t packed = 0xffffffffu;
		uint32_t nIn = 0, nCur = 0;
		if (inDeg <= 2)
		{
			// the usual case, branch-free: both record slots are looked up, an absent neighbour (0xffffffff) matches nothing
			const uint32_t u0 = fl.sh.nbr[tc][0][i][fl.lane], u1 = fl.sh.nbr[tc][1][i][fl.lane];
			uint32_t cu0 = GAF_NONE, cu1 = GAF_NONE, pu0 = GAF_NONE, pu1 = GAF_NONE;
			for (uint32_t j = 0; j < nc; j++)
			{
				const uint32_t id = fl.sh.nodeId[tc][j][fl.lane];
				cu0 = id == u0 ? j : cu0;
				cu1 = id == u1 ? j : cu1;
			}
			for (uint32_t j = 0; j < pNodes; j++)
			{
				const uint32_t id = fl.sh.nodeId[tp][j][fl.lane];
				pu0 = id == u0 ? j : pu0;
				pu1 = id == u1 ? j : pu1;
			}
			const bool have0 = (cu0 & pu0) != GAF_NONE, have1 = (cu1 & pu1) != GAF_NONE;
			// pair 0 is filled first: a lone second neighbour moves down
			const uint32_t a = have0 ? (cu0 | (pu0 << 8)) : (cu1 | (pu1 << 8));
			const uint32_t b2 = (have0 && have1) ? (cu1 | (pu1 << 8)) : 0xffffu;
			packed = (have0 || have1) ? (a | (b2 << 16)) : 0xffffffffu;
			nCur = (cu0 != GAF_NONE ? 1u : 0u) + (cu1 != GAF_NONE ? 1u : 0u);
		}
		else
		{
			const uint32_t node = fl.sh.nodeId[tc][i][fl.lane];
			for (uint32_t e = g.inOff[node], eEnd = g.inOff[node + 1]; e < eEnd; e++)
			{
				const uint32_t u = g.inAdj[e];
				const uint32_t cu = gaf_find<S>(fl.ids(tc), nc, u);
				const uint32_t pu = gaf_find<S>(fl.ids(tp), pNodes, u);
				if (cu == GAF_NONE && pu == GAF_NONE) continue;
				if (nIn >= 2) { ok = false; break; }
				packed = nIn == 0 ? ((packed & 0xffff0000u) | cu | (pu << 8)) : ((packed & 0x0000ffffu) | (cu << 16) | (pu << 24));
				nIn++;
				if (cu != GAF_NONE) nCur++;
			}
		}
		inSlots[(size_t)i * S] = packed;
		pend[(size_t)i * S] = nCur;
		if (nCur == 0) order[(size_t)(ready++) * S] = i;
	}
	GA_SYNCWARP();
	if (!ok) { st.status = GA_ERR_NODE_OVERFLOW; return false; }
	// Kahn over the (at most 16 x 16) in-slot table
	for (uint32_t done = 0; done < ready; done++)
	{
		const uint32_t x = order[(size_t)done * S];
		for (uint32_t y = 0; y < nc; y++)
		{
			const uint32_t p = inSlots[(size_t)y * S];
			const uint32_t hits = ((p & 0xffu) == x ? 1u : 0u) + (((p >> 16) & 0xffu) == x ? 1u : 0u);
			const uint32_t left = pend[(size_t)y * S] - hits;
			pend[(size_t)y * S] = left;
			if (hits != 0 && left == 0) order[(size_t)(ready++) * S] = y;
		}
	}
	if (ready != nc) { st.status = GA_ERR_NODE_OVERFLOW; return false; }   // a cycle: the general kernel replays the reference's work list
	// the nodes' columns lie in the slice's slab (and in the tiny array) in evaluation order: a chain of nodes is then one
	// contiguous run of columns, which the traceback walks without looking anything up (GA_CF_LINK)
	// An order entry also says how the node starts: GAF_ORD_CHAIN = its only band in-neighbour is the node evaluated (and
	// therefore stored) right before it, so its first column is one more word step from that node's last column;
	// GAF_ORD_NOUP = that neighbour is not in the previous band.
	uint32_t col = 0;
	uint32_t prevSlot = GAF_NONE;
	for (uint32_t d = 0; d < nc; d++)
	{
		const uint32_t slot = order[(size_t)d * S];
		fl.sh.csPcs[tc][slot][fl.lane] = (fl.sh.csPcs[tc][slot][fl.lane] & 0xffff0000u) | col;
		col += GA_REC_LEN(fl.sh.lenDeg[tc][slot][fl.lane]);
		const uint32_t ins = inSlots[(size_t)slot * S];
		const bool two = (ins >> 16) != 0xffffu;
		const bool chain = !two && (ins & 0xffu) != GAF_NONE && (ins & 0xffu) == prevSlot;
		order[(size_t)d * S] = slot | (chain ? GAF_ORD_CHAIN : 0u) | (((ins >> 8) & 0xffu) == GAF_NONE ? GAF_ORD_NOUP : 0u);
		prevSlot = slot;
	}
	return true;
}

GA_DEV uint32_t gaf_chunk_base(const uint4& c, uint32_t k)
{
	const uint32_t q = (k >> 4) & 3u;
	const uint32_t w = q == 0 ? c.x : q == 1 ? c.y : q == 2 ? c.z : c.w;
	return (w >> ((k & 15u) * 2u)) & 3u;
}

// Whole stream, forward part.  All lanes of the warp call it (active = this lane holds a stream).
template <int S>
GA_DEV void ga_fast_stream(const ga_graph_view& g, const ga_caps& caps, const GaHmmTables& hmm, const GaUmapSchedule& sch, const GaLaneMem& mem, const GaFastLane<S>& fl, bool active,
	const ga_stream_in* in, const uint32_t* peqAux, int initialBandwidth, int rampBandwidth, uint32_t debugFlags, ga_stream_out* out)
{
	const int LANES = S;   // GA_HDR / GA_HNG
	GaStreamState st;
	st.status = GA_OK;
	st.done = !active;
	st.aux = nullptr;
	st.partLen = active ? in->partLen : 0;
	st.nslices = st.partLen / 64;
	st.startNode = active ? in->startNode : 0;
	st.trimRows = active ? in->trimRows : 0;
	st.prevMin = 0;
	st.hmmC = hmm.startCorrect;
	st.hmmF = hmm.startFalse;
	st.histNodeTop = 0;
	st.slicesPushed = 0;
	st.wordColumns = 0;
	st.cyclicSlices = 0;
	st.rampRedos = 0;
#ifdef GA_PHASE_TIMING
	for (int i = 0; i < 16; i++) st.phase[i] = 0;
#endif
	GA_T0(st);
	uint32_t slicesRun = 0;
	if (active && st.nslices > caps.maxSlices) { st.status = GA_ERR_HIST_OVERFLOW; st.done = true; }
	// initial slice (getInitialSliceOnlyOneNode, GraphAligner.h:2945-2960): the seed node, every column 0
	uint32_t pNodes = 0;
	int tp = 0;
	if (!st.done)
	{
		const GaFastRec r = gaf_load_rec(g, st.startNode);
		const uint32_t len = GA_REC_LEN(r.lenDeg);
		uint32_t nc = 0, ncols = 0;
		if (!gaf_band_add<S>(fl, 0, st, nc, ncols, st.startNode, r.lenDeg, GAF_NOPCS, r.seqChunk, r.in0, r.in1, r.out0, r.out1)) st.done = true;
		else
		{
			fl.sh.nodeMin[0][0][fl.lane] = 0;
			for (uint32_t k = 0; k < len; k++) fl.sh.tiny[0][k][fl.lane] = 0;
			GA_HNG(0, 0) = st.startNode; GA_HNG(0, 1) = 0; GA_HNG(0, 2) = 0; GA_HNG(0, 3) = len; GA_HNG(0, 4) = r.seqChunk;
			pNodes = 1;
			st.histNodeTop = 1;
		}
	}
	// match words of slice 0
	if (!st.done)
	{
		for (int k = 0; k < 4; k++) ga_cp_async8(&fl.sh.eq[0][k][fl.lane], (const uint64_t*)mem.peq + k);
		ga_cp_async4(&fl.sh.aux[0][fl.lane], peqAux);
	}
	int s = 0;
	while (true)
	{
		GA_TLAP(st, 5);
		bool run = !st.done && (uint32_t)s < st.nslices;
		if (!GA_WARP_ANY(run)) break;
		const int tc = tp ^ 1;
		const int eb = s & 1;
		uint32_t ncols = 0;
		int nc = 0;
		const uint32_t nodeOff = st.histNodeTop;
		{
			// slice 0 runs with rampBandwidth as in the reference (rampUntil starts at 0, GraphAligner.h:2612)
			const int bandwidth = s == 0 ? rampBandwidth : initialBandwidth;
			nc = gaf_select_band<S>(g, sch, fl, st, run, bandwidth, tp, pNodes, ncols);
			if (run)
			{
				if (nc <= 0) { if (st.status == GA_OK) st.status = GA_ERR_INTERNAL; st.done = true; run = false; }
				else if (nodeOff + (uint32_t)nc >= caps.histNodes) { st.status = GA_ERR_HIST_OVERFLOW; st.done = true; run = false; }
			}
			if (!run) { nc = 0; ncols = 0; }
		}
		GA_SYNCWARP();
		GA_TLAP(st, 0);
		// the first 32 bases of every band node, into the queue's memory (free until the next selection): they arrive while
		// the band is resolved, and a node start in the column loop reads shared memory only
		for (int i = 0; i < nc; i++) ga_cp_async8(&fl.sh.heap[i][fl.lane], (const uint64_t*)g.seqChunks + (size_t)fl.sh.chunk[tc][i][fl.lane] * 2);
		if (!gaf_resolve<S>(g, fl, st, tc, (uint32_t)nc, run ? pNodes : 0) && run) { st.done = true; run = false; nc = 0; ncols = 0; }
		GA_SYNCWARP();
		GA_TLAP(st, 1);
		// this slice's columns for all lanes of the warp: one chunk of the global history pool
		const uint32_t maxc = GA_WARP_MAX(ncols);
		const uint64_t slabOff64 = GA_POOL_ALLOC(mem.colPoolTop, maxc);
		if (slabOff64 + maxc > caps.warpCols)
		{
			if (run) { st.status = GA_ERR_COL_OVERFLOW; st.done = true; }
			break;   // warp-uniform
		}
		const uint32_t slabOff = (uint32_t)slabOff64;
		GA_TLAP(st, 14);
		// the match words of this slice have arrived; fetch the next slice's
		ga_cp_async_wait();
		if (run && (uint32_t)(s + 1) < st.nslices)
		{
			for (int k = 0; k < 4; k++) ga_cp_async8(&fl.sh.eq[eb ^ 1][k][fl.lane], (const uint64_t*)mem.peq + (size_t)(s + 1) * 4 + k);
			ga_cp_async4(&fl.sh.aux[eb ^ 1][fl.lane], peqAux + (s + 1));
		}
		// ---- the slice's columns: one flat loop --------------------------------------------------------------------
		// Per iteration every lane with columns left computes one column.  The step itself is straight-line code (one basic
		// block: the lone warp of a scheduler lives on instruction-level parallelism); the only branches are the node
		// change (every ~node length iterations per lane) and the rare min-merge with the previous slice's ramp.
		const uint64_t* eqTab = &fl.sh.eq[eb][0][fl.lane];
		const uint32_t auxWord = run ? fl.sh.aux[eb][fl.lane] : 0;
		const uint32_t prevCharCode = auxWord & 7u;
		const bool firstSlice = s == 0;
		const uint32_t* inSlots = fl.scratch();
		const uint32_t* order = fl.scratch() + (size_t)16 * S;
		const int32_t tinyRef = st.prevMin;
		const int32_t INF = 0x3fffffff;
		uint32_t oi = 0;                 // next entry of the evaluation order
		uint32_t slot = 0;               // current node's band slot
		uint32_t kLeft = 0;              // columns of the current node still to do
		uint32_t k = 0;                  // column inside the node
		uint32_t colIdx = 0;             // cs + k: column inside the slice
		uint32_t topIdx = 0;             // pcs + k: the same column in the previous slice
		bool inPrev = false;
		uint32_t prevMask = 0;
		uint64_t seqBits = 0;            // the next (up to 32) bases of the node
		uint32_t chunkIdx = 0;
		uint64_t VP = 0, VN = 0;         // the left neighbour: the column of the previous iteration
		int32_t sbsL = 0, endL = 0;
		uint32_t LsbE = 0;
		int32_t upScore = INF, upRow62 = 0;   // the up-left neighbour: previous iteration's column in the previous slice (INF = none)
		uint32_t lastSlot = GAF_NONE;    // slot whose last column is the left neighbour
		int32_t nodeMin = 0x7fffffff, scoreMax = 0;
		if (!run) oi = (uint32_t)nc;
		for (uint32_t it = 0; it < maxc; it++)
		{
			// (a lane takes up to two columns per pass: the warp is done when no lane has a column left)
			const bool work = kLeft > 0 || oi < (uint32_t)nc;
			if (!GA_WARP_ANY(work)) break;
			if (!work) continue;
			uint32_t isFirst = 0;
			if (kLeft == 0)
			{
				// ---- next node of the evaluation order ----
				if (lastSlot != GAF_NONE) fl.sh.nodeMin[tc][lastSlot][fl.lane] = nodeMin;
				const uint32_t ord = order[(size_t)(oi++) * S];
				slot = ord & 0xffu;
				const uint32_t len = GA_REC_LEN(fl.sh.lenDeg[tc][slot][fl.lane]);
				const uint32_t cp = fl.sh.csPcs[tc][slot][fl.lane];
				const uint32_t cs = cp & 0xffffu, pcs = cp >> 16;
				inPrev = pcs != GAF_NOPCS;
				colIdx = cs;
				topIdx = inPrev ? pcs : 0;
				prevMask = firstSlice ? (inPrev ? 15u : 0u) : ((1u << prevCharCode) & 15u);
				chunkIdx = fl.sh.chunk[tc][slot][fl.lane];
				seqBits = fl.sh.heap[slot][fl.lane];
				k = 0;
				kLeft = len;
				nodeMin = 0x7fffffff;
				// a chain (GAF_ORD_CHAIN): the node's only band in-neighbour is the node evaluated just before it - its last column
				// is the left neighbour in the registers, and its last column of the previous slice (if any) was the previous
				// iteration's up neighbour: the first column is then one more word step of the loop below, and the column is
				// stored right after its neighbour's (GA_CF_LINK)
				if (ord & GAF_ORD_CHAIN)
				{
					isFirst = 1;
					if (ord & GAF_ORD_NOUP) upScore = INF;
				}
				else
				{
					// ---- first column of every other node (GraphAligner.h:1270-1347,1457-1531), cf. ga_calc_node ----
					const uint32_t ins = inSlots[(size_t)slot * S];
					const uint32_t cu0 = ins & 0xffu, pu0 = (ins >> 8) & 0xffu, cu1 = (ins >> 16) & 0xffu, pu1 = ins >> 24;
					const uint32_t base = (uint32_t)seqBits & 3u;
					const uint64_t Eq = eqTab[(size_t)base * S];
					const bool previousEq = ((prevMask >> base) & 1u) != 0;
					const uint32_t oldTiny0 = inPrev ? fl.tinyLd(tp, pcs, tinyRef) : 0;
					// row -1 score of the column and its "exists" flag (forceComponentZeroRow, GraphAligner.h:1916-1989)
					int32_t sbs0 = inPrev ? ga_tiny_score(oldTiny0) : 0x7fffffff;
					uint32_t curCol0 = 0xffffffffu, curCol1 = 0xffffffffu, prevCol0 = 0xffffffffu, prevCol1 = 0xffffffffu;
					uint32_t nIn = 0;
					for (uint32_t e = 0; e < 2; e++)
					{
						const uint32_t cu = e == 0 ? cu0 : cu1, pu = e == 0 ? pu0 : pu1;
						if (cu == GAF_NONE && pu == GAF_NONE) continue;
						uint32_t cc = 0xffffffffu, pc = 0xffffffffu;
						if (cu != GAF_NONE)
						{
							cc = (fl.sh.csPcs[tc][cu][fl.lane] & 0xffffu) + GA_REC_LEN(fl.sh.lenDeg[tc][cu][fl.lane]) - 1;
							const int32_t v = ga_col_load_sbs<S>(mem, slabOff + cc) + 1;
							if (v < sbs0) sbs0 = v;
						}
						if (pu != GAF_NONE)
						{
							pc = (fl.sh.csPcs[tp][pu][fl.lane] & 0xffffu) + GA_REC_LEN(fl.sh.lenDeg[tp][pu][fl.lane]) - 1;
							const int32_t v = ga_tiny_score(fl.tinyLd(tp, pc, tinyRef)) + 1;
							if (v < sbs0) sbs0 = v;
						}
						if (nIn == 0) { curCol0 = cc; prevCol0 = pc; } else { curCol1 = cc; prevCol1 = pc; }
						nIn++;
					}
					const bool sbE0 = inPrev && ga_tiny_score(oldTiny0) == sbs0;
					GaCol c0;
					c0.VP = 0; c0.VN = 0; c0.sbs = 0; c0.scoreEnd = 0;
					const bool single = nIn == 1 && curCol0 != 0xffffffffu;
					uint32_t flags0 = 0;
					if (nIn > 0)
					{
						for (uint32_t e = 0; e < nIn; e++)
						{
							const uint32_t curCol = e == 0 ? curCol0 : curCol1, prevCol = e == 0 ? prevCol0 : prevCol1;
							const bool foundOneUp = prevCol != 0xffffffffu;
							const uint32_t up = foundOneUp ? fl.tinyLd(tp, prevCol, tinyRef) : 0;
							GaCol Ln;
							bool LnsbE;
							uint64_t EqHere = Eq;
							if (curCol != 0xffffffffu)
							{
								Ln = ga_col_load<S>(mem, slabOff + curCol);
								const uint32_t t = fl.tinyLd(tc, curCol, tinyRef);
								Ln.scoreEnd = ga_tiny_score(t);
								LnsbE = (t & 4u) != 0;
							}
							else
							{
								// neighbour only in the previous band: synthetic source column from its end score (GraphAligner.h:1294-1301)
								const int32_t es = ga_tiny_score(up);
								Ln.VP = ~(uint64_t)0;
								Ln.VN = 0;
								Ln.sbs = es;
								Ln.scoreEnd = es + 64;
								LnsbE = true;
								EqHere &= 1;
							}
							uint32_t eq0;
							bool needMerge;
							GaCol cand = ga_next_col(EqHere, Ln, LnsbE, sbE0 && foundOneUp, foundOneUp, previousEq, ga_tiny_row62(up), (single && inPrev) ? ga_tiny_score(oldTiny0) : 0x7fffffff, eq0, needMerge);
							if (single)
							{
								if (!needMerge && curCol + 1 == cs) flags0 = GA_CF_LINK | (eq0 ? GA_CF_EQ0 : 0u);
								if (needMerge) ga_vertical_merge(cand, ga_tiny_score(oldTiny0));
							}
							if (e == 0) c0 = cand;
							else c0 = ga_merge_cols(c0, cand);
						}
						if (!single && inPrev && c0.sbs > ga_tiny_score(oldTiny0)) ga_vertical_merge(c0, ga_tiny_score(oldTiny0));
					}
					else
					{
						// source node (GraphAligner.h:1317-1347,1475-1488); a band node always has a band predecessor or is kept
						if (!inPrev) { st.status = GA_ERR_INTERNAL; oi = (uint32_t)nc; kLeft = 0; continue; }
						const int32_t ps = ga_tiny_score(oldTiny0);
						uint64_t mismatch = 1;
						if (firstSlice) mismatch = (((auxWord >> 4) >> base) & 1u) ? 0 : 1;
						c0.VP = (~(uint64_t)1) | mismatch;
						c0.VN = 0;
						c0.scoreEnd = ps + 63 + (int32_t)mismatch;
						c0.sbs = ps;
					}
					ga_col_store<S>(mem, slabOff + cs, c0, flags0);
					fl.sh.tiny[tc][cs][fl.lane] = (uint16_t)ga_tiny_pack(c0, sbE0);
					nodeMin = c0.scoreEnd;
					if (c0.scoreEnd > scoreMax) scoreMax = c0.scoreEnd;
					VP = c0.VP; VN = c0.VN; sbsL = c0.sbs; endL = c0.scoreEnd;
					LsbE = sbE0 ? 1u : 0u;
					upScore = inPrev ? ga_tiny_score(oldTiny0) : INF;
					upRow62 = ga_tiny_row62(oldTiny0);
					lastSlot = slot;
					seqBits >>= 2;
					k = 1;
					kLeft = len - 1;
					colIdx++;
					topIdx++;
					continue;
				}
			}
			// ---- one column by the word step: columns 1.. of a node (GraphAligner.h:1349-1399,1532-1570) or the first column of a chain ----
			// (a lambda so that a lane can take two columns of a node in one pass of the loop: the second step's shared-memory reads
			// and address arithmetic then overlap the first step's stores - one warp per sub-partition has no other warp to fill its stalls)
			auto wordStep = [&](const uint32_t first) {
				// the next 32 bases of a node longer than that: 8 bytes of its chunks
				if ((k & 31u) == 0 && k > 0) seqBits = *((const uint64_t*)g.seqChunks + (size_t)chunkIdx * 2 + (k >> 5));
				const uint32_t base = (uint32_t)seqBits & 3u;
				seqBits >>= 2;
				uint64_t Eq = eqTab[(size_t)base * S];
				const uint32_t previousEq = (prevMask >> base) & 1u;
				// this column in the previous slice
				const uint32_t topRaw = fl.sh.tiny[tp][topIdx][fl.lane];
				const uint32_t topDec = (uint32_t)tinyRef + (((topRaw >> 3) - (uint32_t)tinyRef) & 0x1fffu);
				const int32_t topScore = inPrev ? (int32_t)topDec : INF;
				const int32_t topRow62 = (int32_t)topDec - (int32_t)(topRaw & 1u) + (int32_t)((topRaw >> 1) & 1u);
				const bool upPresent = upScore != INF;
				// row -1 score = min(left + 1, up-left + 1, previous slice's end score); the flag says the latter attains it
				const int32_t s1 = sbsL + 1;
				const int32_t sbs0 = upScore + 1 < s1 ? upScore + 1 : s1;
				const bool sbE = topScore <= sbs0;
				// ga_next_col, straight-line
				if (!(LsbE && upPresent)) Eq &= ~(uint64_t)1;
				const int32_t dgn = upRow62 + 1 - (int32_t)previousEq;
				int32_t sbsN = (sbE && upPresent && dgn < s1) ? dgn : s1;
				const bool lower = topScore < sbsN;
				const bool legal = topScore >= sbsL - 1;
				sbsN = (lower && legal) ? topScore : sbsN;
				const bool needMerge = lower && !legal;
				const int32_t hin = sbsN - sbsL;
				const uint64_t Xv = Eq | VN;
				if (hin < 0) Eq |= 1;
				const uint32_t eq0 = (uint32_t)Eq & 1u;
				const uint64_t Xh = (((Eq & VP) + VP) ^ VP) | Eq;
				uint64_t Ph = VN | ~(Xh | VP);
				uint64_t Mh = VP & Xh;
				int32_t endN = endL + (int32_t)(Ph >> 63) - (int32_t)(Mh >> 63);
				Ph = (Ph << 1) | (hin > 0 ? 1u : 0u);
				Mh = (Mh << 1) | (hin < 0 ? 1u : 0u);
				uint64_t VPn = Mh | ~(Xv | Ph);
				uint64_t VNn = Ph & Xv;
				uint32_t flags = first ? GA_CF_LINK : GA_CF_PLAIN;
				if (needMerge)
				{
					GaCol c;
					c.VP = VPn; c.VN = VNn; c.sbs = sbsN; c.scoreEnd = endN;
					ga_vertical_merge(c, topScore);
					VPn = c.VP; VNn = c.VN; sbsN = c.sbs; endN = c.scoreEnd;
					flags = 0;
				}
				{
					uint4 ra;
					ra.x = (uint32_t)VPn; ra.y = (uint32_t)(VPn >> 32); ra.z = (uint32_t)VNn; ra.w = (uint32_t)(VNn >> 32);
					const size_t idx = (size_t)(slabOff + colIdx) * S;
					mem.colVV[idx] = ra;
					mem.colS[idx] = (uint32_t)sbsN | flags | ((flags && eq0) ? GA_CF_EQ0 : 0u);
				}
				fl.sh.tiny[tc][colIdx][fl.lane] = (uint16_t)(((uint32_t)endN << 3) | (sbE ? 4u : 0u) | (uint32_t)((VNn >> 62) & 2) | (uint32_t)(VPn >> 63));
				nodeMin = endN < nodeMin ? endN : nodeMin;
				scoreMax = endN > scoreMax ? endN : scoreMax;
				VP = VPn; VN = VNn; sbsL = sbsN; endL = endN;
				LsbE = sbE ? 1u : 0u;
				upScore = topScore;
				upRow62 = topRow62;
				lastSlot = slot;
				k++;
				kLeft--;
				colIdx++;
				topIdx += inPrev ? 1u : 0u;
			};
			wordStep(isFirst);
#pragma unroll
			for (int u = 1; u < GAF_STEPS; u++) { if (kLeft > 0) wordStep(0); }
		}
		const bool failed = st.status != GA_OK;
		GA_TLAP(st, 3);
		if (!run) { s++; continue; }
		if (failed) { st.done = true; s++; continue; }
		if (lastSlot != GAF_NONE) fl.sh.nodeMin[tc][lastSlot][fl.lane] = nodeMin;
		// the 16-bit tiny encoding holds scores up to GA_TINY_SPAN above the reference: beyond that the stream leaves this kernel
		if (scoreMax - tinyRef > GA_TINY_SPAN) { st.status = GA_ERR_COL_OVERFLOW; st.done = true; s++; continue; }
		slicesRun++;
		// ---- slice minimum, HMM step, stop rule (GraphAligner.h:2375,2410-2418,2640-2647; AlignmentCorrectnessEstimation.cpp:71-89) ----
		int32_t minScore = 0x7fffffff;
		for (uint32_t i = 0; i < (uint32_t)nc; i++)
		{
			const int32_t v = fl.sh.nodeMin[tc][i][fl.lane];
			if (v < minScore) minScore = v;
		}
		st.wordColumns += ncols;
		const int32_t m = minScore - st.prevMin;
		if (m < 0 || m > 64) { st.status = GA_ERR_INTERNAL; st.done = true; s++; continue; }
		const double cc = st.hmmC + hmm.c2c, fc = st.hmmF + hmm.f2c;
		const double cf = st.hmmC + hmm.c2f, ff = st.hmmF + hmm.f2f;
		const bool correctFromCorrect = cc >= fc;
		const bool falseFromCorrect = cf >= ff;
		const double hmmC = (cc > fc ? cc : fc) + hmm.correctMul[m];
		const double hmmF = (cf > ff ? cf : ff) + hmm.falseMul[m];
		const bool currentlyCorrect = hmmC > hmmF;
		if (!correctFromCorrect) { st.done = true; s++; continue; }   // stop, slice not recorded
		// the slice is kept: header and node list go to the history
		GA_HDR(s, 0) = slabOff;
		GA_HDR(s, 1) = ncols;
		GA_HDR(s, 2) = nodeOff;
		GA_HDR(s, 3) = (uint32_t)nc;
		GA_HDR(s, 4) = (uint32_t)minScore;
		GA_HDR(s, 5) = (currentlyCorrect ? 1u : 0u) | (falseFromCorrect ? 2u : 0u);
		{
			const uint64_t c = ga_double_to_bits(hmmC), f = ga_double_to_bits(hmmF);
			GA_HDR(s, 6) = (uint32_t)c; GA_HDR(s, 7) = (uint32_t)(c >> 32);
			GA_HDR(s, 8) = (uint32_t)f; GA_HDR(s, 9) = (uint32_t)(f >> 32);
		}
		GA_HDR(s, 10) = 0xffffffffu;   // no cyclic component: the tied minimum cells are ordered at the end (ga_finish_stream)
		GA_HDR(s, 11) = 0;
		for (uint32_t i = 0; i < (uint32_t)nc; i++)
		{
			GA_HNG(nodeOff + i, 0) = fl.sh.nodeId[tc][i][fl.lane];
			GA_HNG(nodeOff + i, 1) = fl.sh.csPcs[tc][i][fl.lane] & 0xffffu;
			GA_HNG(nodeOff + i, 2) = (uint32_t)fl.sh.nodeMin[tc][i][fl.lane];
			GA_HNG(nodeOff + i, 3) = GA_REC_LEN(fl.sh.lenDeg[tc][i][fl.lane]);
			GA_HNG(nodeOff + i, 4) = fl.sh.chunk[tc][i][fl.lane];
		}
		st.hmmC = hmmC;
		st.hmmF = hmmF;
		st.prevMin = minScore;
		st.slicesPushed = (uint32_t)s + 1;
		st.histNodeTop = nodeOff + (uint32_t)nc;
		pNodes = (uint32_t)nc;
		tp = tc;
		s++;
		GA_TLAP(st, 5);
	}
	ga_cp_async_wait();
	ga_finish_stream<S>(g, caps, mem, st, active, slicesRun, debugFlags, out);
}

#endif
