// Traceback kernel body: a warp walks the stored column history of T streams backwards
// (reference getTraceFromTable / getTraceFromTableInner / pickBacktracePredecessor, GraphAligner.h:493-591,894-1021).
//
// The forward pass (ga_core.cuh / ga_fast.cuh) leaves per column {VP, VN} and the row -1 score with three flags.  A backward
// step inside a node needs three facts about the cell: is the left neighbour one lower (horizontal), is the diagonal
// neighbour lower by the mismatch cost (diagonal), else vertical - the reference's order of candidates.  For a column that
// is a plain Myers word step from the column stored before it those facts are bits of the step's own masks: H (= Ph, rows
// one above the left column) and D0 (rows whose diagonal delta is 0), with EQ the match word of the column's base; the
// diagonal is taken where D0 == EQ.  The masks are not stored; they are re-derived from the left neighbour's {VP, VN}.
// Everything else - first columns of nodes with several in-neighbours, slice borders, min-merged columns - takes the
// general path, which evaluates the reference's candidates from cell values.
//
// A traceback is a chain of dependent reads into gigabytes of history; what the kernel is built around is never to have a
// lane wait alone:
//   * WINDOWS.  The 32 lanes of the warp together fetch, for each of the warp's T streams, the W = 32 P history columns to
//     the left of where that stream's walk stands (asynchronous copies, all streams' requests in flight at once), then
//     turn them into the two decision masks per column {H, D0 xnor EQ} in shared memory.  The forward pass stores a chain
//     of nodes as one contiguous run of columns, so a window spans node borders.
//   * WALK.  Each of the T stream lanes then walks its own window out of shared memory: two bit tests per move, no global
//     reads, until it leaves the window or meets a column / row that needs the general path.
//   * GENERAL STEPS are deferred: a lane that needs one parks until every lane of the warp has parked, then all of them
//     take their general step together, so that its chain of global reads is paid once per round and not once per lane.
//   * SLICES IN LOCK STEP.  The streams of a warp walk the same slice at the same time: header, node list and match words
//     of the next slice are fetched by all lanes together.
// The same source compiles for the host (oracle/hostsim): lanes become loops over per-stream state.
#ifndef GA_TRACE_CUH
#define GA_TRACE_CUH
#include "ga_core.cuh"

struct GaTraceMem
{
	size_t S;                    // lane interleave of the forward launch (streams per warp); every pointer is offset by the lane
	const uint32_t* hdr;
	const uint32_t* histNode;
	const uint4* colVV;
	const uint32_t* colS;
	const uint4* peq;            // this stream's match words, two 16-byte halves per slice
	uint32_t* moves;             // temporary trace record of this stream (interleaved like the rest)
	uint32_t* pathNodes;
	uint32_t* runs;
	uint32_t maxMoves, maxPathNodes, maxRuns;
};

#define GA_TR_NODES 32      /* band nodes of a slice kept in shared memory (a larger band is walked by general steps only) */
#define GA_TR_NONE 0xffffffffu
#define GA_TR_CROSS 8       /* nodes a lane may leave in one walk phase */

// one warp's shared memory, [..][stream lane]
template <int T, int P>
struct GaTraceShared
{
	uint4 wHD[P * 32][T];                // window entry e = history column (window top - e): {H, D} of rows 0..31, {H, D} of rows 32..63 (staging: the left column's {VP, VN})
	uint32_t wF[P * 32][T];              //   1 = the column is the first one of its node (staging: score word)
	uint32_t wB[P * 32][T];              //   staging: the 16 bases around the column's base
	uint32_t wM[P * 32][T];              //   staging: bit 0 entry valid, bit 1 offset in node > 0, bits 2..5 offset & 15
	uint64_t peq[4][T];                  // match words of the slice the lane is in
	uint32_t nodeId[GA_TR_NODES][T], nodeCs[GA_TR_NODES][T], nodeLen[GA_TR_NODES][T], nodeChunk[GA_TR_NODES][T];   // that slice's band
	uint32_t nodeBefore[GA_TR_NODES][T];  // the band node whose columns end where this one's begin (GA_TR_NODES: none)
	uint32_t upId[GA_TR_NODES][T], upCs[GA_TR_NODES][T], upLen[GA_TR_NODES][T], upChunk[GA_TR_NODES][T];   // the band of the slice above
	// window requests of the stream lanes to the warp
	uint32_t reqTop[T];                  // history column wanted as entry 0 (GA_TR_NONE: no request)
	uint32_t reqSlot[T], reqOff[T];      //   its band slot and offset in the node
	uint32_t reqLo[T];                   //   first column of the slice's slab (nothing below it belongs to the slice)
	uint32_t reqLink[T];                 //   GaTraceState::linkFast
	uint32_t cross[GA_TR_CROSS][T];      // walk: row | diagonal << 8 of the moves that left a node
	const uint4* colVV[T];               // the streams' history pointers (GaTraceMem)
	const uint32_t* colS[T];
};

#if defined(__CUDACC__)
GA_DEV void ga_tr_cp16(void* dstShared, const void* src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dstShared)), "l"(src));
}
GA_DEV void ga_tr_cp4(void* dstShared, const void* src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((uint32_t)__cvta_generic_to_shared(dstShared)), "l"(src));
}
GA_DEV void ga_tr_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// per-stream phases run on the stream's lane, cooperative phases on all 32 lanes; the host runs both as loops
#define GA_TR_EACH(l) if (const uint32_t l = lane; l < (uint32_t)T)
#define GA_TR_COOP(i) if (const uint32_t i = lane; true)
#define GA_TR_LANE(l) lanes[0]
#define GA_TR_ANY(field) __any_sync(0xffffffffu, lane < (uint32_t)T && lanes[0].t.field)
#else
GA_DEV void ga_tr_cp16(void* dst, const void* src) { memcpy(dst, src, 16); }
GA_DEV void ga_tr_cp4(void* dst, const void* src) { memcpy(dst, src, 4); }
GA_DEV void ga_tr_cp_wait() {}
#define GA_TR_EACH(l) for (uint32_t l = 0; l < (uint32_t)T; l++)
#define GA_TR_COOP(i) for (uint32_t i = 0; i < 32; i++)
#define GA_TR_LANE(l) lanes[l]
template <typename F> static inline bool ga_tr_any_host(int n, F f) { bool r = false; for (int l = 0; l < n; l++) r = r || f(l); return r; }
#define GA_TR_ANY(field) ga_tr_any_host(T, [&](int l_) { return (bool)lanes[l_].t.field; })
#endif

// host emulation: event counters of the walk (GA_TRACE_COUNTS=1 prints them when the library unloads)
#if !defined(__CUDACC__)
struct GaTrCounts
{
	unsigned long long rounds = 0, fillRounds = 0, fills = 0, genRounds = 0, gens = 0, fastSteps = 0, linkSteps = 0, slices = 0, genRow0 = 0, genFlag = 0, genUncached = 0, genOther = 0;
	~GaTrCounts()
	{
		if (getenv("GA_TRACE_COUNTS")) fprintf(stderr, "[ga trace counts] slices %llu rounds %llu fillRounds %llu fills %llu genRounds %llu gens %llu (row0 %llu flag %llu uncached %llu other %llu) fast %llu link %llu\n",
			slices, rounds, fillRounds, fills, genRounds, gens, genRow0, genFlag, genUncached, genOther, fastSteps, linkSteps);
	}
};
static GaTrCounts g_trCounts;
#define GA_TR_COUNT(x) g_trCounts.x++
#else
#define GA_TR_COUNT(x)
#endif

#define GA_TR_HDR(s, f) tm.hdr[(size_t)((size_t)(s) * GA_HDR_WORDS + (f)) * tm.S]
#define GA_TR_HN(i, f) tm.histNode[(size_t)((size_t)(i) * GA_HN_WORDS + (f)) * tm.S]

struct GaTrCol
{
	uint64_t VP, VN;
	int32_t sbs;
};

GA_DEV GaTrCol ga_tr_col(const GaTraceMem& tm, uint32_t col)
{
	const uint4 a = tm.colVV[(size_t)col * tm.S];
	GaTrCol c;
	c.VP = (uint64_t)a.x | ((uint64_t)a.y << 32);
	c.VN = (uint64_t)a.z | ((uint64_t)a.w << 32);
	c.sbs = (int32_t)(tm.colS[(size_t)col * tm.S] & GA_CF_SCORE_MASK);
	return c;
}

GA_DEV int32_t ga_tr_end_score(const GaTraceMem& tm, uint32_t col)
{
	const GaTrCol c = ga_tr_col(tm, col);
	return c.sbs + (int32_t)GA_POPC(c.VP) - (int32_t)GA_POPC(c.VN);
}

GA_DEV int ga_tr_find(const GaTraceMem& tm, uint32_t nodeOff, uint32_t nNodes, uint32_t node)
{
	int r = -1;
	for (uint32_t i = 0; i < nNodes; i++)
	{
		if (GA_TR_HN(nodeOff + i, 0) == node) r = (int)i;
	}
	return r;
}

// value of (node, off) at `row` of slice s, or `maxv` when the node is not in that slice's band
// (getValueOrMax, GraphAligner.h:2008-2017).  s == -1 is the initial slice: seed node = 0, else maxv.
// ALT (-B ramp runs, ga_run_stream<.., RAMP>): seen from the slice below, a slice is the instance the reference kept as a sqrt
// checkpoint (GA_HF_ALT: header index in word 1), which after a ramp redo need not be the instance the walk goes through.
template <bool ALT = false>
GA_DEV int32_t ga_tr_value(const GaTraceMem& tm, uint32_t startNode, int s, uint32_t node, uint32_t off, int row, int32_t maxv)
{
	if (s < 0) return node == startNode ? 0 : maxv;
	if (ALT && (GA_TR_HDR(s, 5) & GA_HF_ALT)) s = (int)GA_TR_HDR(s, 1);
	const uint32_t nodeOff = GA_TR_HDR(s, 2), nNodes = GA_TR_HDR(s, 3);
	const int slot = ga_tr_find(tm, nodeOff, nNodes, node);
	if (slot < 0) return maxv;
	const GaTrCol c = ga_tr_col(tm, GA_TR_HDR(s, 0) + GA_TR_HN(nodeOff + slot, 1) + off);
	return ga_col_value(c.VP, c.VN, c.sbs, row);
}

// per-stream walk state
struct GaTraceState
{
	bool walking;
	int32_t status;
	uint32_t nMoves, nPath, nRuns, skipped, curWord;
	bool runOpen;
	uint32_t runNode, runLastOff, runLastRow;
	int s, row;
	uint32_t node, off;
	// the slice the lane is in
	uint32_t sNodeOff, sNodes, sSlab;
	bool sCached;
	// the slice above (fetched when the lane enters a slice: the walk will need its last row, then its tables)
	int upSlice;                 // slice whose band the up table holds (-1: none)
	uint32_t upNodeOff, upNodes, upSlab;
	bool upCached;
	bool upAlt;                  // ALT: the up table holds the checkpoint instance of the slice above, not the one to walk through
	int hintSlice, hintSlot;     // band slot (or -1) of node hintNode in slice hintSlice, as found in that slice's cached table
	uint32_t hintNode;
	bool locate;                 // node changed without a link: look it up in the slice's band again
	uint32_t slot, colBase, nodeLen, nodeChunk;
	uint32_t winTop;             // history column of entry 0 of the lane's window (GA_TR_NONE: nothing loaded)
	bool linkFast;               // the stream's score is below the read length: no cell of the trace can equal the "not in the band" value
	bool needFill, needGen;      // what the lane asks the warp for at the end of a walk phase
	uint4 chunk;                 // general path: 64 bases of the node, starting at base chunkBase
	uint32_t chunkBase;
};

// everything one stream lane owns
struct GaTraceLane
{
	GaTraceMem tm;
	const ga_stream_in* in;
	GaTraceState t;
};

// base `off` of the node the walk is in, from the 64 bases held in registers
GA_DEV uint32_t ga_tr_node_base(const ga_graph_view& g, GaTraceState& t, uint32_t off)
{
	if ((off & ~63u) != t.chunkBase)
	{
		t.chunkBase = off & ~63u;
		t.chunk = *((const uint4*)g.seqChunks + t.nodeChunk + (off >> 6));
	}
	const uint32_t q = (off >> 4) & 3u;
	const uint32_t w = q == 0 ? t.chunk.x : (q == 1 ? t.chunk.y : (q == 2 ? t.chunk.z : t.chunk.w));
	return (w >> ((off & 15u) * 2u)) & 3u;
}

// end score (row 63) of column `off` of `node` in the slice above the lane's, or maxv when the node is not in that band
// (getValueOrMax on the previous slice, GraphAligner.h:2008-2017); the band comes from the up table in shared memory
template <int T, int P, bool ALT = false>
GA_DEV int32_t ga_tr_up_end(const GaTraceMem& tm, GaTraceShared<T, P>& sh, uint32_t lane, GaTraceState& t, uint32_t startNode, uint32_t node, uint32_t off, int32_t maxv)
{
	if (t.s <= 0) return node == startNode ? 0 : maxv;
	if (!t.upCached || t.upSlice != t.s - 1) return ga_tr_value<ALT>(tm, startNode, t.s - 1, node, off, 63, maxv);
	int slot = t.hintSlot;
	if (t.hintSlice != t.s - 1 || t.hintNode != node)
	{
		// (slice, node) -> slot is remembered: the walk asks for two columns of a node, then usually moves up into it
		slot = -1;
		for (uint32_t i = 0; i < t.upNodes; i++) if (sh.upId[i][lane] == node) slot = (int)i;
		// (a slot of the checkpoint instance's band says nothing about the instance to walk through: the hint is kept under a
		// slice number no walk is ever in)
		t.hintSlice = (ALT && t.upAlt) ? -3 : t.s - 1;
		t.hintNode = node;
		t.hintSlot = slot;
	}
	if (slot < 0) return maxv;
	return ga_tr_end_score(tm, t.upSlab + sh.upCs[slot][lane] + off);
}

#define GA_TR_EMIT(mv) \
	{ \
		t.curWord |= (uint32_t)(mv) << ((t.nMoves & 15) * 2); \
		t.nMoves++; \
		if ((t.nMoves & 15) == 0) { tm.moves[(size_t)((t.nMoves >> 4) - 1) * tm.S] = t.curWord; t.curWord = 0; } \
	}

// Trimmed-tail bookkeeping, once per position the walk stands on: a run (maximal stretch of positions on one node) opens
// at the first untrimmed position seen on a node; positions in the trimmed tail are only counted.
GA_DEV void ga_tr_account(const ga_stream_in& in, GaTraceState& t)
{
	if (t.runOpen) return;
	const uint32_t j = (uint32_t)t.s * 64u + (uint32_t)t.row;
	if (j < in.trimRows) { t.runOpen = true; t.runNode = t.node; t.runLastOff = t.off; t.runLastRow = j; }
	else t.skipped++;
}

// (slice, node) -> band slot, column base, node length; false = the walk failed
template <int T, int P>
GA_DEV bool ga_tr_locate(const GaTraceMem& tm, GaTraceShared<T, P>& sh, uint32_t lane, GaTraceState& t)
{
	int found = -1;
	if (t.sCached)
	{
		if (t.hintSlice == t.s && t.hintNode == t.node) found = t.hintSlot;
		else for (uint32_t i = 0; i < t.sNodes; i++) if (sh.nodeId[i][lane] == t.node) found = (int)i;
	}
	else found = ga_tr_find(tm, t.sNodeOff, t.sNodes, t.node);
	if (found < 0) { t.status = GA_ERR_TRACE; t.walking = false; return false; }
	t.slot = (uint32_t)found;
	if (t.sCached) { t.colBase = t.sSlab + sh.nodeCs[t.slot][lane]; t.nodeLen = sh.nodeLen[t.slot][lane]; t.nodeChunk = sh.nodeChunk[t.slot][lane]; }
	else { t.colBase = t.sSlab + GA_TR_HN(t.sNodeOff + t.slot, 1); t.nodeLen = GA_TR_HN(t.sNodeOff + t.slot, 3); t.nodeChunk = GA_TR_HN(t.sNodeOff + t.slot, 4); }
	t.chunkBase = 0xffffffffu;
	t.locate = false;
	return true;
}

// WALK phase of one stream lane inside slice sw: steps out of the window until the lane needs the warp (a window, a
// general step) or has left the slice.
//
// A window entry holds, per 32-row half, the words {H, D}: bit r of H = the move at row r is horizontal, of D = it is
// diagonal (D is only set where H is not), neither = vertical; both set in every row = the column is not a plain word step
// from the column stored before it (general step).  Which neighbour "the column before" is depends on the offset in the
// node: the left neighbour in the node (GA_CF_PLAIN), or for a node's first column the last column of its only band
// in-neighbour (GA_CF_LINK) - the same three candidates then (in-neighbour horizontal, in-neighbour diagonal, vertical:
// GraphAligner.h:501-533 with one neighbour).
template <int T, int P>
GA_DEV void ga_tr_walk(GaTraceShared<T, P>& sh, uint32_t lane, GaTraceLane& L, int sw)
{
	const GaTraceMem& tm = L.tm;
	GaTraceState& t = L.t;
	t.needFill = t.needGen = false;
	sh.reqTop[lane] = GA_TR_NONE;
	if (!(t.walking && t.s == sw)) return;
	if (t.locate && !ga_tr_locate<T, P>(tm, sh, lane, t)) return;
	if (!t.sCached) { t.needGen = true; GA_TR_COUNT(genUncached); return; }
	const ga_stream_in& in = *L.in;
	const uint32_t W = (uint32_t)(P * 32);
	while (true)
	{
		// row 0 is the slice border (the candidates lie in the slice above): general step.  The moves of one pass through
		// the loop body below (at most W + 64) must fit the record.
		if (t.row <= 0 || t.nMoves + W + 64 >= tm.maxMoves) { t.needGen = true; GA_TR_COUNT(genRow0); return; }
		const uint32_t x = t.colBase + t.off;
		uint32_t e = t.winTop - x;
		if (t.winTop == GA_TR_NONE || e >= W)
		{
			t.needFill = true;
			GA_TR_COUNT(fills);
			t.winTop = x;
			sh.reqTop[lane] = x;
			sh.reqSlot[lane] = t.slot;
			sh.reqOff[lane] = t.off;
			sh.reqLo[lane] = t.sSlab;
			sh.reqLink[lane] = t.linkFast ? 1u : 0u;
			return;
		}
		if (!t.runOpen)
		{
			// ---- the trimmed tail (all lanes of a warp are in it together, at the start of the trace): one step with the
			// bookkeeping of a position that may open the stream's first run ----
			const uint4 hd = sh.wHD[e][lane];
			const uint32_t hw = t.row >= 32 ? hd.z : hd.x, dw = t.row >= 32 ? hd.w : hd.y;
			const uint32_t hbit = (hw >> (t.row & 31)) & 1u, dbit = (dw >> (t.row & 31)) & 1u;
			if (hbit & dbit) { t.needGen = true; GA_TR_COUNT(genFlag); return; }
			// the reference's order of candidates: horizontal, diagonal, vertical
			const uint32_t mv = hbit ? (uint32_t)GA_MOVE_H : (dbit ? (uint32_t)GA_MOVE_D : (uint32_t)GA_MOVE_V);
			const bool first = t.off == 0;
			uint32_t nslot = 0;
			if (first)
			{
				if (t.nPath >= tm.maxPathNodes || t.nRuns >= tm.maxRuns) { t.needGen = true; GA_TR_COUNT(genFlag); return; }
				if (mv != GA_MOVE_V)
				{
					nslot = sh.nodeBefore[t.slot][lane];
					if (nslot >= GA_TR_NODES) { t.needGen = true; GA_TR_COUNT(genOther); return; }
				}
			}
			ga_tr_account(in, t);
			GA_TR_EMIT(mv);
			if (first) { GA_TR_COUNT(linkSteps); } else { GA_TR_COUNT(fastSteps); }
			if (!first) { if (mv != GA_MOVE_V) t.off--; }
			else if (mv != GA_MOVE_V)
			{
				// leaving the node: a run just opened on this very position is closed again
				if (t.runOpen)
				{
					uint32_t* r = tm.runs + (size_t)(t.nRuns * GA_RUN_WORDS) * tm.S;
					r[0] = t.runNode; r[tm.S] = 0; r[2 * tm.S] = t.runLastOff; r[3 * tm.S] = (uint32_t)t.s * 64u + (uint32_t)t.row; r[4 * tm.S] = t.runLastRow;
					t.nRuns++;
					t.runOpen = false;
				}
				const uint32_t nlen = sh.nodeLen[nslot][lane];
				t.slot = nslot;
				t.node = sh.nodeId[nslot][lane];
				t.nodeChunk = sh.nodeChunk[nslot][lane];
				t.chunkBase = 0xffffffffu;
				t.nodeLen = nlen;
				t.off = nlen - 1;
				t.colBase -= nlen;
				tm.pathNodes[(size_t)t.nPath * tm.S] = t.node;
				t.nPath++;
			}
			t.row -= (mv != GA_MOVE_H) ? 1 : 0;
			continue;
		}
		// ---- steady state, the run of the node the walk is in is open and every position from here on is untrimmed (rows
		// only decrease): one flat loop, a move per iteration, all lanes of the warp in step.  Leaving a node (horizontal or
		// diagonal move out of a first column) only notes the row; the records are written after the loop.
		const uint32_t e0 = e;
		uint32_t row = (uint32_t)t.row, nM = t.nMoves, cw = t.curWord, nCross = 0;
		uint32_t* mp = tm.moves + (size_t)(nM >> 4) * tm.S;
		bool stop = false;
		do
		{
			const uint32_t* hp = (const uint32_t*)&sh.wHD[e][lane] + ((row >> 5) << 1);
			const uint32_t hbit = (hp[0] >> (row & 31u)) & 1u, dbit = (hp[1] >> (row & 31u)) & 1u;
			const uint32_t first = sh.wF[e][lane];
			if (hbit & dbit) { stop = true; break; }
			const uint32_t adv = hbit | dbit;
			if (first & adv) { sh.cross[nCross][lane] = row | (dbit << 8); nCross++; GA_TR_COUNT(linkSteps); }
			const uint32_t mv = 2u - 2u * hbit - dbit;   // GA_MOVE_H = 0, GA_MOVE_D = 1, GA_MOVE_V = 2
			cw |= mv << ((nM & 15u) * 2u);
			nM++;
			if ((nM & 15u) == 0) { *mp = cw; mp += tm.S; cw = 0; }
			e += adv;
			row -= 1u - hbit;
			GA_TR_COUNT(fastSteps);
		} while (row != 0 && e < W && nCross < GA_TR_CROSS);
		t.row = (int)row; t.nMoves = nM; t.curWord = cw;
		// the nodes left on the way: close each one's run, cross into the node stored before it, open that node's run
		uint32_t rem = e - e0;
		for (uint32_t k = 0; k < nCross; k++)
		{
			const uint32_t cr = sh.cross[k][lane];
			const uint32_t crow = cr & 0xffu, isD = cr >> 8;
			if (t.nRuns >= tm.maxRuns || t.nPath >= tm.maxPathNodes) { t.status = GA_ERR_TRACE_OVERFLOW; t.walking = false; return; }
			uint32_t* r = tm.runs + (size_t)(t.nRuns * GA_RUN_WORDS) * tm.S;
			r[0] = t.runNode; r[tm.S] = 0; r[2 * tm.S] = t.runLastOff; r[3 * tm.S] = (uint32_t)t.s * 64u + crow; r[4 * tm.S] = t.runLastRow;
			t.nRuns++;
			rem -= t.off + 1;
			const uint32_t nslot = sh.nodeBefore[t.slot][lane];
			const uint32_t nlen = sh.nodeLen[nslot][lane];
			t.slot = nslot;
			t.node = sh.nodeId[nslot][lane];
			t.nodeLen = nlen;
			t.off = nlen - 1;
			t.colBase -= nlen;
			tm.pathNodes[(size_t)t.nPath * tm.S] = t.node;
			t.nPath++;
			t.runNode = t.node; t.runLastOff = t.off; t.runLastRow = (uint32_t)t.s * 64u + crow - isD;
		}
		if (nCross) { t.nodeChunk = sh.nodeChunk[t.slot][lane]; t.chunkBase = 0xffffffffu; }
		t.off -= rem;
		if (stop) { t.needGen = true; GA_TR_COUNT(genFlag); return; }
	}
}

// GENERAL step of one stream lane (node starts, slice borders, merged columns): the reference's candidates from the stored
// columns' cell values.  Leaves t.s decremented when the walk went up a slice.
template <int T, int P, bool ALT = false>
GA_DEV void ga_tr_general(const ga_graph_view& g, GaTraceShared<T, P>& sh, uint32_t lane, GaTraceLane& L)
{
	const GaTraceMem& tm = L.tm;
	GaTraceState& t = L.t;
	const ga_stream_in& in = *L.in;
	const int32_t maxv = (int32_t)in.partLen;
	const uint32_t startNode = in.startNode;
#define GA_TR_FAIL(code) { t.status = (code); t.walking = false; return; }
	ga_tr_account(in, t);
	GA_TR_COUNT(gens);
	uint32_t move = 4;
	uint32_t nnode = t.node, noff = t.off;
	const int s = t.s, row = t.row;
	const uint32_t off = t.off, node = t.node;
	{
		const GaTrCol cur = ga_tr_col(tm, t.colBase + off);
		const int32_t here = ga_col_value(cur.VP, cur.VN, cur.sbs, row);
		const uint32_t base = ga_tr_node_base(g, t, off);
		const int32_t match = (int32_t)((sh.peq[base][lane] >> row) & 1);
		const int32_t diagWant = here - 1 + match;
		const bool firstRow = s == 0 && row == 0;
		if (firstRow && node == startNode && (here == 0 || here == 1))
		{
			move = GA_MOVE_END;   // GraphAligner.h:500
		}
		else if (off > 0)
		{
			const GaTrCol left = ga_tr_col(tm, t.colBase + off - 1);
			const int32_t hs = ga_col_value(left.VP, left.VN, left.sbs, row);
			int32_t ds, us;
			if (row > 0)
			{
				ds = hs - (int32_t)((left.VP >> row) & 1) + (int32_t)((left.VN >> row) & 1);
				us = here - (int32_t)((cur.VP >> row) & 1) + (int32_t)((cur.VN >> row) & 1);
			}
			else if (s == 0)
			{
				ds = us = node == startNode ? 0 : maxv;   // the initial slice: seed node all zero
			}
			else
			{
				// row 63 of the slice above; a node's columns are contiguous there too
				ds = ga_tr_up_end<T, P, ALT>(tm, sh, lane, t, startNode, node, off - 1, maxv);
				us = ga_tr_up_end<T, P, ALT>(tm, sh, lane, t, startNode, node, off, maxv);
			}
			if (hs == here - 1) { move = GA_MOVE_H; noff = off - 1; }
			else if (ds == diagWant) { move = GA_MOVE_D; noff = off - 1; }
			else if (us == here - 1) { move = GA_MOVE_V; }
		}
		else
		{
			// first column of a node: in-neighbours in inNeighbors order, horizontal before diagonal (GraphAligner.h:501-533)
			for (uint32_t e = g.inOff[node], eEnd = g.inOff[node + 1]; e < eEnd; e++)
			{
				const uint32_t u = g.inAdj[e];
				// the neighbour's last column in this slice: from the shared-memory node list when there is one
				int uslot = -1;
				uint32_t uoff;
				if (t.sCached)
				{
					for (uint32_t i = 0; i < t.sNodes; i++) if (sh.nodeId[i][lane] == u) uslot = (int)i;
					uoff = uslot >= 0 ? sh.nodeLen[uslot][lane] - 1 : (uint32_t)(g.nodeStart[u + 1] - g.nodeStart[u]) - 1;
				}
				else
				{
					uslot = ga_tr_find(tm, t.sNodeOff, t.sNodes, u);
					uoff = (uint32_t)(g.nodeStart[u + 1] - g.nodeStart[u]) - 1;
				}
				GaTrCol uc;
				uc.VP = uc.VN = 0; uc.sbs = 0;
				if (uslot >= 0) uc = ga_tr_col(tm, t.sSlab + (t.sCached ? sh.nodeCs[uslot][lane] : GA_TR_HN(t.sNodeOff + uslot, 1)) + uoff);
				const int32_t hs = uslot >= 0 ? ga_col_value(uc.VP, uc.VN, uc.sbs, row) : maxv;
				if (hs == here - 1) { move = GA_MOVE_H; nnode = u; noff = uoff; break; }
				int32_t ds;
				if (row == 0) ds = ga_tr_up_end<T, P, ALT>(tm, sh, lane, t, startNode, u, uoff, maxv);
				else ds = uslot >= 0 ? ga_col_value(uc.VP, uc.VN, uc.sbs, row - 1) : maxv;
				if (ds == diagWant) { move = GA_MOVE_D; nnode = u; noff = uoff; break; }
			}
			if (move == 4)
			{
				int32_t us;
				if (row > 0) us = here - (int32_t)((cur.VP >> row) & 1) + (int32_t)((cur.VN >> row) & 1);
				else us = ga_tr_up_end<T, P, ALT>(tm, sh, lane, t, startNode, node, off, maxv);
				if (us == here - 1) { move = GA_MOVE_V; }
			}
		}
		// any step into row -1 ends the trace; that last position is popped again (GraphAligner.h:949-951)
		if (firstRow && (move == GA_MOVE_D || move == GA_MOVE_V)) move = GA_MOVE_END;
		if (move == 4) GA_TR_FAIL(GA_ERR_TRACE);   // reference: assert(false); std::abort()
	}
	if (t.nMoves >= tm.maxMoves) GA_TR_FAIL(GA_ERR_TRACE_OVERFLOW);
	GA_TR_EMIT(move);
	// leaving the node (or ending): close the open run; its first position is the one we stand on
	if (t.runOpen && (move == GA_MOVE_END || nnode != node))
	{
		if (t.nRuns >= tm.maxRuns) GA_TR_FAIL(GA_ERR_TRACE_OVERFLOW);
		uint32_t* r = tm.runs + (size_t)(t.nRuns * GA_RUN_WORDS) * tm.S;
		r[0] = t.runNode; r[tm.S] = off; r[2 * tm.S] = t.runLastOff; r[3 * tm.S] = (uint32_t)s * 64u + (uint32_t)row; r[4 * tm.S] = t.runLastRow;
		t.nRuns++;
		t.runOpen = false;
	}
	if (move == GA_MOVE_END) { t.walking = false; return; }
	if (move != GA_MOVE_V && off == 0)
	{
		if (t.nPath >= tm.maxPathNodes) GA_TR_FAIL(GA_ERR_TRACE_OVERFLOW);
		tm.pathNodes[(size_t)t.nPath * tm.S] = nnode;
		t.nPath++;
		t.locate = true;
	}
	if (move != GA_MOVE_H)
	{
		t.row--;
		if (t.row < 0)
		{
			t.row = 63;
			t.s--;
			t.locate = true;
		}
	}
	t.node = nnode;
	t.off = noff;
#undef GA_TR_FAIL
}

// WINDOW fill, first half: entry e of stream j's request - where its column, its left neighbour and its base are, and the
// asynchronous copies of all three into the window's staging words
template <int T, int P>
GA_DEV void ga_tr_fill_issue(const ga_graph_view& g, GaTraceShared<T, P>& sh, size_t S, uint32_t j, uint32_t e)
{
	const uint32_t top = sh.reqTop[j], lo = sh.reqLo[j];
	uint32_t meta = 0;
	// the column (top - e) and the one stored before it must lie in the slice's slab
	if (top - lo >= e + 1)
	{
		const uint32_t c = top - e;
		// node and offset of the column: back along the chain of nodes stored one after the other
		uint32_t slot = sh.reqSlot[j];
		int32_t off = (int32_t)sh.reqOff[j] - (int32_t)e;
		while (off < 0 && slot < GA_TR_NODES)
		{
			slot = sh.nodeBefore[slot][j];
			if (slot < GA_TR_NODES) off += (int32_t)sh.nodeLen[slot][j];
		}
		// a first column is walked through only into the node stored before it, and only when no cell of the trace can equal the
		// reference's "not in the band" value (linkFast)
		if (slot < GA_TR_NODES && off == 0 && (sh.reqLink[j] == 0 || sh.nodeBefore[slot][j] >= GA_TR_NODES)) slot = GA_TR_NODES;
		if (slot < GA_TR_NODES)
		{
			ga_tr_cp16(&sh.wHD[e][j], sh.colVV[j] + (size_t)(c - 1) * S);
			ga_tr_cp4(&sh.wF[e][j], sh.colS[j] + (size_t)c * S);
			ga_tr_cp4(&sh.wB[e][j], g.seqChunks + (size_t)sh.nodeChunk[slot][j] * 4 + ((uint32_t)off >> 4));
			meta = 1u | (off > 0 ? 2u : 0u) | (((uint32_t)off & 15u) << 2);
		}
	}
	sh.wM[e][j] = meta;
}

// second half: the staged words become the two decision masks of the column (see ga_tr_walk)
template <int T, int P>
GA_DEV void ga_tr_fill_masks(GaTraceShared<T, P>& sh, uint32_t j, uint32_t e)
{
	const uint32_t meta = sh.wM[e][j];
	uint4 hd = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
	uint32_t first = 0;
	if (meta & 1u)
	{
		const uint32_t sw = sh.wF[e][j];
		if (sw & ((meta & 2u) ? GA_CF_PLAIN : GA_CF_LINK))
		{
			const uint4 lv = sh.wHD[e][j];
			const uint64_t lvp = (uint64_t)lv.x | ((uint64_t)lv.y << 32), lvn = (uint64_t)lv.z | ((uint64_t)lv.w << 32);
			const uint32_t base = (sh.wB[e][j] >> (((meta >> 2) & 15u) * 2u)) & 3u;
			const uint64_t EQ = sh.peq[base][j];
			// the horizontal half of ga_next_col with the match bit of row 0 as the forward pass used it
			const uint64_t Eq = (EQ & ~(uint64_t)1) | ((sw & GA_CF_EQ0) ? 1u : 0u);
			const uint64_t Xh = (((Eq & lvp) + lvp) ^ lvp) | Eq;
			const uint64_t H = lvn | ~(Xh | lvp);
			const uint64_t D0 = Xh | lvn;
			// the diagonal delta 1 - D0 must equal the mismatch cost 1 - EQ
			const uint64_t D = ~(D0 ^ EQ) & ~H;
			hd = make_uint4((uint32_t)H, (uint32_t)D, (uint32_t)(H >> 32), (uint32_t)(D >> 32));
			first = (meta & 2u) ? 0u : 1u;
		}
	}
	sh.wHD[e][j] = hd;
	sh.wF[e][j] = first;
}

// Entering slice sw (one stream lane): its header, match words and node list (already here when the lane came down from
// the slice below: it was that slice's up table), then the tables of the slice above.
template <int T, int P, bool ALT = false>
GA_DEV void ga_tr_enter_slice(GaTraceShared<T, P>& sh, uint32_t lane, GaTraceLane& L, int sw)
{
	const GaTraceMem& tm = L.tm;
	GaTraceState& t = L.t;
	for (int k = 0; k < 2; k++)
	{
		const uint4 q = tm.peq[(size_t)sw * 2 + k];
		sh.peq[k * 2][lane] = (uint64_t)q.x | ((uint64_t)q.y << 32);
		sh.peq[k * 2 + 1][lane] = (uint64_t)q.z | ((uint64_t)q.w << 32);
	}
	const bool fromUp = t.upSlice == sw && !(ALT && t.upAlt);
	if (fromUp) { t.sSlab = t.upSlab; t.sNodeOff = t.upNodeOff; t.sNodes = t.upNodes; t.sCached = t.upCached; }
	else
	{
		t.sSlab = GA_TR_HDR(sw, 0);
		t.sNodeOff = GA_TR_HDR(sw, 2);
		t.sNodes = GA_TR_HDR(sw, 3);
		t.sCached = t.sNodes <= GA_TR_NODES;
	}
	uint32_t upSlab = 0, upNodeOff = 0, upNodes = 0;
	if (sw > 0)
	{
		uint32_t uh = (uint32_t)(sw - 1);
		if (ALT)
		{
			t.upAlt = (GA_TR_HDR(uh, 5) & GA_HF_ALT) != 0;
			if (t.upAlt) uh = GA_TR_HDR(uh, 1);
		}
		upSlab = GA_TR_HDR(uh, 0); upNodeOff = GA_TR_HDR(uh, 2); upNodes = GA_TR_HDR(uh, 3);
	}
	if (t.sCached)
	{
		for (uint32_t i = 0; i < t.sNodes; i++)
		{
			if (fromUp)
			{
				sh.nodeId[i][lane] = sh.upId[i][lane]; sh.nodeCs[i][lane] = sh.upCs[i][lane]; sh.nodeLen[i][lane] = sh.upLen[i][lane]; sh.nodeChunk[i][lane] = sh.upChunk[i][lane];
			}
			else
			{
				ga_tr_cp4(&sh.nodeId[i][lane], &GA_TR_HN(t.sNodeOff + i, 0));
				ga_tr_cp4(&sh.nodeCs[i][lane], &GA_TR_HN(t.sNodeOff + i, 1));
				ga_tr_cp4(&sh.nodeLen[i][lane], &GA_TR_HN(t.sNodeOff + i, 3));
				ga_tr_cp4(&sh.nodeChunk[i][lane], &GA_TR_HN(t.sNodeOff + i, 4));
			}
		}
	}
	t.upSlice = -1;
	if (sw > 0)
	{
		t.upSlab = upSlab; t.upNodeOff = upNodeOff; t.upNodes = upNodes;
		t.upCached = upNodes <= GA_TR_NODES;
		t.upSlice = sw - 1;
		if (t.upCached)
		{
			for (uint32_t i = 0; i < upNodes; i++)
			{
				ga_tr_cp4(&sh.upId[i][lane], &GA_TR_HN(upNodeOff + i, 0));
				ga_tr_cp4(&sh.upCs[i][lane], &GA_TR_HN(upNodeOff + i, 1));
				ga_tr_cp4(&sh.upLen[i][lane], &GA_TR_HN(upNodeOff + i, 3));
				ga_tr_cp4(&sh.upChunk[i][lane], &GA_TR_HN(upNodeOff + i, 4));
			}
		}
	}
	ga_tr_cp_wait();
	if (t.sCached)
	{
		for (uint32_t i = 0; i < t.sNodes; i++)
		{
			const uint32_t myCs = sh.nodeCs[i][lane];
			uint32_t before = GA_TR_NODES;
			for (uint32_t k = 0; k < t.sNodes; k++)
			{
				const uint32_t kl = sh.nodeLen[k][lane];
				before = (sh.nodeCs[k][lane] + kl == myCs && kl > 0) ? k : before;
			}
			sh.nodeBefore[i][lane] = before;
		}
	}
	t.winTop = GA_TR_NONE;
	t.locate = true;
}

// Walks the T streams of a warp.  Device: called by all 32 lanes, lanes[0] is the calling lane's stream (lanes >= T have
// none and only help with the windows).  Host: lanes[0..T).
template <int T, int P, bool ALT = false>
GA_DEV void ga_trace_warp(const ga_graph_view& g, GaTraceShared<T, P>& sh, uint32_t lane, GaTraceLane* lanes, size_t S)
{
	(void)lane;
	uint32_t sTopMine = 0;
	GA_TR_EACH(l)
	{
		GaTraceLane& L = GA_TR_LANE(l);
		sh.colVV[l] = L.tm.colVV;
		sh.colS[l] = L.tm.colS;
		sh.reqTop[l] = GA_TR_NONE;
		if (L.t.walking && (uint32_t)(L.t.s + 1) > sTopMine) sTopMine = (uint32_t)(L.t.s + 1);
	}
#if defined(__CUDACC__)
	const int sTop = (int)__reduce_max_sync(0xffffffffu, sTopMine);
#else
	const int sTop = (int)sTopMine;
#endif
	GA_SYNCWARP();
	for (int sw = sTop - 1; sw >= 0; sw--)
	{
		GA_TR_EACH(l)
		{
			GaTraceLane& L = GA_TR_LANE(l);
			if (L.t.walking && L.t.s == sw) ga_tr_enter_slice<T, P, ALT>(sh, l, L, sw);
		}
		GA_SYNCWARP();
		GA_TR_COUNT(slices);
		while (true)
		{
			GA_TR_EACH(l) ga_tr_walk<T, P>(sh, l, GA_TR_LANE(l), sw);
			GA_SYNCWARP();
			const bool anyFill = GA_TR_ANY(needFill), anyGen = GA_TR_ANY(needGen);
			if (!anyFill && !anyGen) break;
			GA_TR_COUNT(rounds);
			if (anyGen) { GA_TR_COUNT(genRounds); }
			if (anyFill)
			{
				GA_TR_COUNT(fillRounds);
				// all requests' copies in flight, then all masks
				for (uint32_t j = 0; j < (uint32_t)T; j++)
				{
					if (sh.reqTop[j] == GA_TR_NONE) continue;
					for (uint32_t p = 0; p < (uint32_t)P; p++) GA_TR_COOP(i) ga_tr_fill_issue<T, P>(g, sh, S, j, p * 32 + i);
				}
				ga_tr_cp_wait();
				GA_SYNCWARP();
				for (uint32_t j = 0; j < (uint32_t)T; j++)
				{
					if (sh.reqTop[j] == GA_TR_NONE) continue;
					for (uint32_t p = 0; p < (uint32_t)P; p++) GA_TR_COOP(i) ga_tr_fill_masks<T, P>(sh, j, p * 32 + i);
				}
				GA_SYNCWARP();
			}
			if (anyGen)
			{
				GA_TR_EACH(l)
				{
					GaTraceLane& L = GA_TR_LANE(l);
					if (L.t.needGen) ga_tr_general<T, P, ALT>(g, sh, l, L);
				}
			}
			GA_SYNCWARP();
		}
	}
	GA_TR_EACH(l)
	{
		GaTraceLane& L = GA_TR_LANE(l);
		if (L.t.nMoves & 15) L.tm.moves[(size_t)(L.t.nMoves >> 4) * L.tm.S] = L.t.curWord;
	}
}

// initial state of a stream's walk (doTrace = the forward pass left a trace start)
GA_DEV void ga_trace_init(GaTraceLane& L, bool doTrace, int nSlices, uint32_t node, uint32_t off, int32_t score)
{
	GaTraceState& t = L.t;
	t.walking = doTrace;
	t.status = GA_OK;
	t.nMoves = t.nPath = t.nRuns = t.skipped = t.curWord = 0;
	t.runOpen = false;
	t.runNode = t.runLastOff = t.runLastRow = 0;
	t.s = nSlices - 1;
	t.row = 63;
	t.node = node;
	t.off = off;
	t.sNodeOff = t.sNodes = t.sSlab = 0;
	t.sCached = false;
	t.locate = true;
	t.slot = t.colBase = t.nodeLen = t.nodeChunk = 0;
	t.winTop = GA_TR_NONE;
	t.linkFast = doTrace && score < (int32_t)L.in->partLen;
	t.needFill = t.needGen = false;
	t.chunk = make_uint4(0, 0, 0, 0);
	t.chunkBase = 0xffffffffu;
	t.upSlice = -1;
	t.upNodeOff = t.upNodes = t.upSlab = 0;
	t.upCached = false;
	t.upAlt = false;
	t.hintSlice = -2;
	t.hintSlot = -1;
	t.hintNode = 0;
}

// ---- mapping records on the device (GA_SRC_SOLO streams) ---------------------------------------------------------
// traceToAlignment (GraphAligner.h:782-847) for a forward stream whose rows are not shifted: mapping j = run (nRuns-1-j) of the
// backward-ordered run list; only the first mapping has an offset, every mapping but the last covers its last base too
// (from_length = lastOff - firstOff + 1), to_length counts the read rows since the previous mapping's last row.
// false = a run lies on one of the graph's two dummy nodes (the host's mappingRange handles that: keep the runs).
GA_DEV bool ga_tr_runs_mappable(const GaTraceMem& tm, uint32_t nRuns, uint32_t nNodes)
{
	for (uint32_t k = 0; k < nRuns; k++)
	{
		const uint32_t node = tm.runs[(size_t)(k * GA_RUN_WORDS) * tm.S];
		if (node == 0 || node + 1 == nNodes) return false;
	}
	return nRuns > 0;
}

GA_DEV GaDeviceMapping ga_tr_mapping(const ga_graph_view& g, const uint32_t* runs, size_t S, uint32_t nRuns, uint32_t j)
{
	const uint32_t* r = runs + (size_t)((nRuns - 1 - j) * GA_RUN_WORDS) * S;
	const uint32_t node = r[0], firstOff = r[S], lastOff = r[2 * S], firstJ = r[3 * S], lastJ = r[4 * S];
	const uint32_t beforeJ = j == 0 ? firstJ : runs[(size_t)((nRuns - j) * GA_RUN_WORDS + 4) * S];
	const long long idRev = g.nodeIdRev[node];
	GaDeviceMapping m;
	m.node_id = idRev >> 1;
	m.offset = j == 0 ? firstOff : 0;
	m.rank = j;
	m.from_length = (int32_t)(lastOff - firstOff) + (j + 1 == nRuns ? 0 : 1);
	m.to_length = (int32_t)(lastJ - beforeJ);
	m.read_start = firstJ;
	m.is_reverse = (uint32_t)(idRev & 1);
	return m;
}

// positions = the start cell plus one per move except the terminating one, minus the trimmed tail
GA_DEV uint32_t ga_trace_positions(const GaTraceState& t) { return t.nMoves > t.skipped ? t.nMoves - t.skipped : 0; }

#endif
