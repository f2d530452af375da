// Traceback kernel body: one lane walks one stream's stored column history backwards
// (reference getTraceFromTable / getTraceFromTableInner / pickBacktracePredecessor, GraphAligner.h:493-591,894-1021).
//
// The forward pass (ga_core.cuh / ga_fast.cuh) leaves per column {VP, VN} and the row -1 score with three flags.  A backward
// step inside a node needs three facts about the cell: is the left neighbour one lower (horizontal), is the diagonal
// neighbour lower by the mismatch cost (diagonal), else vertical - the reference's order of candidates.  For a column that
// is a plain Myers word step from the column stored before it those facts are bits of the step's own masks: H (= Ph, rows
// one above the left column) and D0 (rows whose diagonal delta is 0), with EQ the match word of the column's base.  They are
// not stored; the walk re-derives them from the left neighbour's {VP, VN} when it enters a column.  Everything else - first
// columns of nodes with several in-neighbours, slice borders, min-merged columns, the trimmed tail - takes the general path,
// which evaluates the reference's candidates from cell values.
//
// What keeps the lanes of a warp fed and together:
//   * a traceback is a chain of dependent reads into a history of gigabytes.  Each lane keeps two windows of 32 consecutive
//     history columns in shared memory, filled by asynchronous copies (cp.async): the walk reads the window it is in while
//     the next one (the 31 columns to the left) is already on its way.  The forward pass stores a chain of nodes as one
//     contiguous run of columns (ga_fast.cuh), so a window usually spans node borders;
//   * the 32 streams of a warp walk the same slice at the same time (outer loop over slices, a vote per step): the border
//     between two slices - header, node list, match words of the next slice, the cell values of the row above - is then
//     crossed by all lanes together instead of one lane at a time.
// The same source compiles for the host with one lane (oracle/hostsim).
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

#define GA_TR_WIN 32        /* columns per window; consecutive windows overlap by one column (the left neighbour of the last entry) */
#define GA_TR_NODES 16      /* band nodes of the current slice kept in shared memory (larger bands are searched in global memory) */

// one warp's shared memory, [..][lane]
template <int L>
struct GaTraceShared
{
	uint4 wVV[2][GA_TR_WIN][L];          // two windows of {VP, VN}: entry e of a window = history column (window top - e)
	uint32_t wS[2][GA_TR_WIN][L];        //   and of the score words (row -1 score | GA_CF_* flags)
	uint64_t peq[4][L];                  // match words of the slice the lane is in
	uint32_t nodeId[GA_TR_NODES][L], nodeCs[GA_TR_NODES][L], nodeLen[GA_TR_NODES][L], nodeChunk[GA_TR_NODES][L];   // that slice's band
	uint4 nodeSeq[GA_TR_NODES][L];        //   and the first 64 bases of each of its nodes
	uint32_t nodeBefore[GA_TR_NODES][L];  // the band node whose columns end where this one's begin (GA_TR_NODES: none)
	uint32_t upId[GA_TR_NODES][L], upCs[GA_TR_NODES][L], upLen[GA_TR_NODES][L], upChunk[GA_TR_NODES][L];   // the band of the slice above
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
#else
GA_DEV void ga_tr_cp16(void* dst, const void* src) { memcpy(dst, src, 16); }
GA_DEV void ga_tr_cp4(void* dst, const void* src) { memcpy(dst, src, 4); }
GA_DEV void ga_tr_cp_wait() {}
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
GA_DEV int32_t ga_tr_value(const GaTraceMem& tm, uint32_t startNode, int s, uint32_t node, uint32_t off, int row, int32_t maxv)
{
	if (s < 0) return node == startNode ? 0 : maxv;
	const uint32_t nodeOff = GA_TR_HDR(s, 2), nNodes = GA_TR_HDR(s, 3);
	const int slot = ga_tr_find(tm, nodeOff, nNodes, node);
	if (slot < 0) return maxv;
	const GaTrCol c = ga_tr_col(tm, GA_TR_HDR(s, 0) + GA_TR_HN(nodeOff + slot, 1) + off);
	return ga_col_value(c.VP, c.VN, c.sbs, row);
}

// base `off` of a node whose sequence starts at chunk `chunk` (ga_graph_view::seqChunks)
GA_DEV uint32_t ga_tr_base(const ga_graph_view& g, uint32_t chunk, uint32_t off)
{
	const uint32_t w = g.seqChunks[(size_t)chunk * 4 + (off >> 4)];
	return (w >> ((off & 15u) * 2u)) & 3u;
}

// per-lane walk state
struct GaTraceState
{
	bool walking;
	int32_t status;
	uint32_t nMoves, nPath, nRuns, skipped, curWord;
	bool runOpen;
	uint32_t runNode, runLastOff, runLastRow;
	int s, row;
	uint32_t node, off;
	int32_t here;
	bool haveHere;
	// the slice the lane is in
	uint32_t sNodeOff, sNodes, sSlab;
	bool sCached;
	// the slice above (fetched when the lane enters a slice: the walk will need its last row, then its tables)
	int upSlice;                 // slice whose band the up table holds (-1: none)
	uint32_t upNodeOff, upNodes, upSlab;
	bool upCached;
	bool locate;                 // node changed without a link: look it up in the slice's band again
	uint32_t slot, colBase, nodeLen, nodeChunk;
	// windows
	int cur;                     // window the walk is in (buffer 0 or 1)
	int64_t topCur, topNext;     // history column of entry 0 of that window / of the other one (-1: nothing loaded / requested)
	uint4 chunk;                 // 64 bases of the node, starting at base chunkBase
	uint32_t chunkBase;
	// masks of the column the walk stands on
	int64_t maskCol;
	uint64_t mH, mD0, mEQ;
	uint32_t mOk;
};

template <int L>
GA_DEV void ga_tr_issue_window(const GaTraceMem& tm, GaTraceShared<L>& sh, uint32_t lane, int b, int64_t top)
{
	for (int e = 0; e < GA_TR_WIN; e++)
	{
		const int64_t idx = top - e;
		if (idx < 0) break;
		ga_tr_cp16(&sh.wVV[b][e][lane], tm.colVV + (size_t)idx * tm.S);
		ga_tr_cp4(&sh.wS[b][e][lane], tm.colS + (size_t)idx * tm.S);
	}
}

// makes history column x (and its left neighbour) available in the current window; returns its entry
template <int L>
GA_DEV int ga_tr_window(const GaTraceMem& tm, GaTraceShared<L>& sh, uint32_t lane, GaTraceState& t, int64_t x)
{
	const int64_t e = t.topCur >= 0 ? t.topCur - x : -1;
	if (e >= 0 && e < GA_TR_WIN - 1) return (int)e;
	if (e == GA_TR_WIN - 1 && t.topNext == x)
	{
		// walked to the end of the window: the next one (requested when this one became current) starts at this column
		ga_tr_cp_wait();
		t.cur ^= 1;
	}
	else
	{
		// somewhere else (new slice, a node that is not stored next to its neighbour): fetch and wait
		ga_tr_cp_wait();
		ga_tr_issue_window<L>(tm, sh, lane, t.cur, x);
		ga_tr_cp_wait();
	}
	t.topCur = x;
	// request the window to the left while this one is walked
	const int64_t nextTop = x - (GA_TR_WIN - 1);
	if (nextTop >= 0) ga_tr_issue_window<L>(tm, sh, lane, t.cur ^ 1, nextTop);
	t.topNext = nextTop >= 0 ? nextTop : -1;
	return 0;
}

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
template <int L>
GA_DEV int32_t ga_tr_up_end(const GaTraceMem& tm, GaTraceShared<L>& sh, uint32_t lane, const GaTraceState& t, uint32_t startNode, uint32_t node, uint32_t off, int32_t maxv)
{
	if (t.s <= 0) return node == startNode ? 0 : maxv;
	if (!t.upCached || t.upSlice != t.s - 1) return ga_tr_value(tm, startNode, t.s - 1, node, off, 63, maxv);
	int slot = -1;
	for (uint32_t i = 0; i < t.upNodes; i++) if (sh.upId[i][lane] == node) slot = (int)i;
	if (slot < 0) return maxv;
	return ga_tr_end_score(tm, t.upSlab + sh.upCs[slot][lane] + off);
}

// One step (one move) of one lane's walk inside slice t.s.  Leaves t.s decremented when the walk went up a slice.
template <int L>
GA_DEV void ga_trace_step(const ga_graph_view& g, const GaTraceMem& tm, GaTraceShared<L>& sh, uint32_t lane, const ga_stream_in& in, GaTraceState& t)
{
	const int32_t maxv = (int32_t)in.partLen;
	const uint32_t startNode = in.startNode;
#define GA_TR_EMIT(mv) \
	{ \
		t.curWord |= (uint32_t)(mv) << ((t.nMoves & 15) * 2); \
		t.nMoves++; \
		if ((t.nMoves & 15) == 0) { tm.moves[(size_t)((t.nMoves >> 4) - 1) * tm.S] = t.curWord; t.curWord = 0; } \
	}
#define GA_TR_FAIL(code) { t.status = (code); t.walking = false; return; }
	if (t.locate)
	{
		int found = -1;
		if (t.sCached)
		{
			for (uint32_t i = 0; i < t.sNodes; i++) if (sh.nodeId[i][lane] == t.node) found = (int)i;
		}
		else found = ga_tr_find(tm, t.sNodeOff, t.sNodes, t.node);
		if (found < 0) GA_TR_FAIL(GA_ERR_TRACE);
		t.slot = (uint32_t)found;
		if (t.sCached) { t.colBase = t.sSlab + sh.nodeCs[t.slot][lane]; t.nodeLen = sh.nodeLen[t.slot][lane]; t.nodeChunk = sh.nodeChunk[t.slot][lane]; }
		else { t.colBase = t.sSlab + GA_TR_HN(t.sNodeOff + t.slot, 1); t.nodeLen = GA_TR_HN(t.sNodeOff + t.slot, 3); t.nodeChunk = GA_TR_HN(t.sNodeOff + t.slot, 4); }
		if (t.sCached) { t.chunk = sh.nodeSeq[t.slot][lane]; t.chunkBase = 0; } else t.chunkBase = 0xffffffffu;
		if (!t.haveHere)
		{
			const GaTrCol c = ga_tr_col(tm, t.colBase + t.off);
			t.here = ga_col_value(c.VP, c.VN, c.sbs, t.row);
			t.haveHere = true;
		}
		t.locate = false;
	}
	if (!t.runOpen)
	{
		// still inside the trimmed tail, or a run was just closed: open one at the first untrimmed position
		const uint32_t j = (uint32_t)t.s * 64u + (uint32_t)t.row;
		if (j < in.trimRows) { t.runOpen = true; t.runNode = t.node; t.runLastOff = t.off; t.runLastRow = j; }
		else t.skipped++;
	}
	const int row = t.row;
	if (t.runOpen && row > 0 && t.nMoves < tm.maxMoves)
	{
		const int64_t x = (int64_t)t.colBase + t.off;
		if (t.maskCol != x)
		{
			// ---- entering a column: the masks of the word step that produced it, from the left neighbour in the window ----
			const int e = ga_tr_window<L>(tm, sh, lane, t, x);
			const uint32_t sw = sh.wS[t.cur][e][lane];
			t.mOk = 0;
			if (sw & (t.off > 0 ? GA_CF_PLAIN : GA_CF_LINK))
			{
				const uint4 lv = sh.wVV[t.cur][e + 1][lane];
				const uint64_t lvp = (uint64_t)lv.x | ((uint64_t)lv.y << 32), lvn = (uint64_t)lv.z | ((uint64_t)lv.w << 32);
				const uint64_t EQ = sh.peq[ga_tr_node_base(g, t, t.off)][lane];
				// the horizontal half of ga_next_col with the match bit of row 0 as the forward pass used it
				const uint64_t Eq = (EQ & ~(uint64_t)1) | ((sw & GA_CF_EQ0) ? 1u : 0u);
				const uint64_t Xh = (((Eq & lvp) + lvp) ^ lvp) | Eq;
				t.mH = lvn | ~(Xh | lvp);
				t.mD0 = Xh | lvn;
				t.mEQ = EQ;
				t.mOk = t.off > 0 ? 1u : 2u;
			}
			t.maskCol = x;
		}
		// ---- fast step: inside the node, inside the slice, on a plain word-step column (nine steps in ten).  The three
		// candidates of pickBacktracePredecessor reduce to three bit tests, taken in the reference's order: horizontal,
		// diagonal, vertical.  Nothing here can fail or change node / slice / run.
		if ((t.mOk & 1u) && t.off > 0)
		{
			const uint32_t hbit = (uint32_t)(t.mH >> row) & 1u;
			const uint32_t d0 = (uint32_t)(t.mD0 >> row) & 1u;
			const uint32_t eq = (uint32_t)(t.mEQ >> row) & 1u;
			// diagonal delta = 1 - D0 must equal the mismatch cost 1 - EQ
			const uint32_t mv = hbit ? (uint32_t)GA_MOVE_H : (d0 == eq ? (uint32_t)GA_MOVE_D : (uint32_t)GA_MOVE_V);
			t.here = t.here - 1 + (int32_t)((mv == GA_MOVE_D) ? d0 : 0u);
			GA_TR_EMIT(mv);
			if (mv != GA_MOVE_V) t.off--;
			t.row -= (mv != GA_MOVE_H) ? 1 : 0;
			return;
		}
		// ---- link step: first column of a node whose only band in-neighbour is in this slice and stored right before it.
		// The column is a word step from that neighbour's last column, so the same three bit tests apply (in-neighbour
		// horizontal, in-neighbour diagonal, vertical: GraphAligner.h:501-533 with one neighbour).
		// here < maxv: an in-neighbour outside the band reads as maxv in the reference and must not be able to match.
		if ((t.mOk & 2u) && t.off == 0 && t.here < maxv && t.nPath < tm.maxPathNodes && t.nRuns < tm.maxRuns)
		{
			const uint32_t hbit = (uint32_t)(t.mH >> row) & 1u;
			const uint32_t d0 = (uint32_t)(t.mD0 >> row) & 1u;
			const uint32_t eq = (uint32_t)(t.mEQ >> row) & 1u;
			const uint32_t mv = hbit ? (uint32_t)GA_MOVE_H : (d0 == eq ? (uint32_t)GA_MOVE_D : (uint32_t)GA_MOVE_V);
			if (mv != GA_MOVE_V)
			{
				// the neighbour = the band node whose columns end right before this node's first one
				uint32_t nslot = 0, nlen = 0;
				if (t.sCached)
				{
					nslot = sh.nodeBefore[t.slot][lane];
					if (nslot < GA_TR_NODES) nlen = sh.nodeLen[nslot][lane];
				}
				else
				{
					const uint32_t myCs = t.colBase - t.sSlab;
					for (uint32_t i = 0; i < t.sNodes; i++)
					{
						const uint32_t ics = GA_TR_HN(t.sNodeOff + i, 1), ilen = GA_TR_HN(t.sNodeOff + i, 3);
						if (ics + ilen == myCs && ilen > 0) { nslot = i; nlen = ilen; }
					}
				}
				if (nlen == 0) GA_TR_FAIL(GA_ERR_TRACE);
				t.here = t.here - 1 + (int32_t)((mv == GA_MOVE_D) ? d0 : 0u);
				GA_TR_EMIT(mv);
				// leaving the node: close the run on the position we stand on, cross into the neighbour
				uint32_t* r = tm.runs + (size_t)(t.nRuns * GA_RUN_WORDS) * tm.S;
				r[0] = t.runNode; r[tm.S] = 0; r[2 * tm.S] = t.runLastOff; r[3 * tm.S] = (uint32_t)t.s * 64u + (uint32_t)row; r[4 * tm.S] = t.runLastRow;
				t.nRuns++;
				t.runOpen = false;
				t.slot = nslot;
				if (t.sCached) { t.node = sh.nodeId[nslot][lane]; t.nodeChunk = sh.nodeChunk[nslot][lane]; }
				else { t.node = GA_TR_HN(t.sNodeOff + nslot, 0); t.nodeChunk = GA_TR_HN(t.sNodeOff + nslot, 4); }
				if (t.sCached) { t.chunk = sh.nodeSeq[nslot][lane]; t.chunkBase = 0; } else t.chunkBase = 0xffffffffu;
				t.nodeLen = nlen;
				t.off = nlen - 1;
				t.colBase -= nlen;
				tm.pathNodes[(size_t)t.nPath * tm.S] = t.node;
				t.nPath++;
			}
			else
			{
				t.here = t.here - 1;
				GA_TR_EMIT(mv);
			}
			t.row -= (mv != GA_MOVE_H) ? 1 : 0;
			return;
		}
	}
	// ---- general path (node starts, slice borders, merged columns, trimmed tail): the reference's candidates from
	// the stored columns' cell values
	uint32_t move = 4;
	uint32_t nnode = t.node, noff = t.off;
	int32_t nhere = 0;
	const int s = t.s;
	const uint32_t off = t.off, node = t.node;
	const int32_t here = t.here;
	{
		const GaTrCol cur = ga_tr_col(tm, t.colBase + off);
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
				ds = ga_tr_up_end<L>(tm, sh, lane, t, startNode, node, off - 1, maxv);
				us = ga_tr_up_end<L>(tm, sh, lane, t, startNode, node, off, maxv);
			}
			if (hs == here - 1) { move = GA_MOVE_H; noff = off - 1; nhere = hs; }
			else if (ds == diagWant) { move = GA_MOVE_D; noff = off - 1; nhere = ds; }
			else if (us == here - 1) { move = GA_MOVE_V; nhere = us; }
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
				if (hs == here - 1) { move = GA_MOVE_H; nnode = u; noff = uoff; nhere = hs; break; }
				int32_t ds;
				if (row == 0) ds = ga_tr_up_end<L>(tm, sh, lane, t, startNode, u, uoff, maxv);
				else ds = uslot >= 0 ? ga_col_value(uc.VP, uc.VN, uc.sbs, row - 1) : maxv;
				if (ds == diagWant) { move = GA_MOVE_D; nnode = u; noff = uoff; nhere = ds; break; }
			}
			if (move == 4)
			{
				int32_t us;
				if (row > 0) us = here - (int32_t)((cur.VP >> row) & 1) + (int32_t)((cur.VN >> row) & 1);
				else us = ga_tr_up_end<L>(tm, sh, lane, t, startNode, node, off, maxv);
				if (us == here - 1) { move = GA_MOVE_V; nhere = us; }
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
	t.here = nhere;
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
#undef GA_TR_EMIT
#undef GA_TR_FAIL
}

// Walks the streams of a warp: called by all lanes (doTrace = this lane has a trace to walk).
template <int L>
GA_DEV void ga_trace_warp(const ga_graph_view& g, const GaTraceMem& tm, GaTraceShared<L>& sh, uint32_t lane, bool doTrace, const ga_stream_in* in, int nSlices, uint32_t node, uint32_t off,
	int32_t& statusOut, uint32_t& nMovesOut, uint32_t& nPathOut, uint32_t& nRunsOut, uint32_t& nPosOut)
{
	GaTraceState t;
	t.walking = doTrace;
	t.status = GA_OK;
	t.nMoves = t.nPath = t.nRuns = t.skipped = t.curWord = 0;
	t.runOpen = false;
	t.runNode = t.runLastOff = t.runLastRow = 0;
	t.s = nSlices - 1;
	t.row = 63;
	t.node = node;
	t.off = off;
	t.here = 0;
	t.haveHere = false;
	t.sNodeOff = t.sNodes = t.sSlab = 0;
	t.sCached = false;
	t.locate = true;
	t.slot = t.colBase = t.nodeLen = t.nodeChunk = 0;
	t.cur = 0;
	t.topCur = t.topNext = -1;
	t.chunk = make_uint4(0, 0, 0, 0);
	t.maskCol = -1;
	t.mH = t.mD0 = t.mEQ = 0;
	t.mOk = 0;
	t.upSlice = -1;
	t.upNodeOff = t.upNodes = t.upSlab = 0;
	t.upCached = false;
	t.chunkBase = 0xffffffffu;
	const int sTop = (int)GA_WARP_MAX(doTrace ? (uint32_t)nSlices : 0u);
	for (int sw = sTop - 1; sw >= 0; sw--)
	{
		const bool mine = t.walking && t.s == sw;
		if (mine)
		{
			// ---- entering a slice: its header, match words and node list (already here when the lane came down from the
			// slice below: it was that slice's up table), then the tables of the slice above ----
			for (int k = 0; k < 2; k++)
			{
				const uint4 q = tm.peq[(size_t)sw * 2 + k];
				sh.peq[k * 2][lane] = (uint64_t)q.x | ((uint64_t)q.y << 32);
				sh.peq[k * 2 + 1][lane] = (uint64_t)q.z | ((uint64_t)q.w << 32);
			}
			const bool fromUp = t.upSlice == sw;
			if (fromUp) { t.sSlab = t.upSlab; t.sNodeOff = t.upNodeOff; t.sNodes = t.upNodes; t.sCached = t.upCached; }
			else
			{
				t.sSlab = GA_TR_HDR(sw, 0);
				t.sNodeOff = GA_TR_HDR(sw, 2);
				t.sNodes = GA_TR_HDR(sw, 3);
				t.sCached = t.sNodes <= GA_TR_NODES;
			}
			uint32_t upSlab = 0, upNodeOff = 0, upNodes = 0;
			if (sw > 0) { upSlab = GA_TR_HDR(sw - 1, 0); upNodeOff = GA_TR_HDR(sw - 1, 2); upNodes = GA_TR_HDR(sw - 1, 3); }
			ga_tr_cp_wait();   // nothing of the old slice's windows is needed any more
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
				if (!fromUp) ga_tr_cp_wait();
				// the nodes' bases (needed per column) come in while the up table does
				for (uint32_t i = 0; i < t.sNodes; i++) ga_tr_cp16(&sh.nodeSeq[i][lane], (const uint4*)g.seqChunks + sh.nodeChunk[i][lane]);
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
			t.topCur = t.topNext = -1;   // the waits above covered the window requests; both are stale anyway
			if (t.sCached)
			{
				for (uint32_t i = 0; i < t.sNodes; i++)
				{
					const uint32_t myCs = sh.nodeCs[i][lane];
					uint32_t before = GA_TR_NODES;
					for (uint32_t j = 0; j < t.sNodes; j++)
					{
						const uint32_t jl = sh.nodeLen[j][lane];
						before = (sh.nodeCs[j][lane] + jl == myCs && jl > 0) ? j : before;
					}
					sh.nodeBefore[i][lane] = before;
				}
			}
			t.locate = true;
			t.maskCol = -1;
		}
		GA_SYNCWARP();
		while (true)
		{
			// inside the slice: every lane on its own (a vote per step keeps the warp together)
			while (true)
			{
				const bool go = t.walking && t.s == sw && !(t.row == 0 && sw > 0);
				if (!GA_WARP_ANY(go)) break;
				if (go) ga_trace_step<L>(g, tm, sh, lane, *in, t);
			}
			// at the slice's first row: the step that may leave the slice reads the last row of the slice above - all lanes
			// that got here take it together, so that their reads overlap
			const bool cross = t.walking && t.s == sw && t.row == 0 && sw > 0;
			if (!GA_WARP_ANY(cross)) break;
			if (cross) ga_trace_step<L>(g, tm, sh, lane, *in, t);
			GA_SYNCWARP();
		}
	}
	ga_tr_cp_wait();
	if (doTrace && (t.nMoves & 15)) tm.moves[(size_t)(t.nMoves >> 4) * tm.S] = t.curWord;
	statusOut = t.status;
	nMovesOut = t.nMoves;
	nPathOut = t.nPath;
	nRunsOut = t.nRuns;
	// positions = the start cell plus one per move except the terminating one, minus the trimmed tail
	nPosOut = t.nMoves > t.skipped ? t.nMoves - t.skipped : 0;
}

#endif
