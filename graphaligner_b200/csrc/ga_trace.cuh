// Traceback kernel body: a group of G lanes walks one stream's stored column history backwards
// (reference getTraceFromTable / getTraceFromTableInner / pickBacktracePredecessor, GraphAligner.h:493-591,894-1021).
//
// The forward pass (ga_core.cuh) leaves per column {VP, VN} and the row -1 score with three flags.  A backward step
// inside a node needs three facts about the cell: is the left neighbour one lower (horizontal), is the diagonal
// neighbour lower by the mismatch cost (diagonal), else vertical - the reference's order of candidates.  For a column that
// is a plain Myers word step from its left neighbour those facts are bits of the step's own masks: H (= Ph, rows one above
// the left column) and D0 (rows whose diagonal delta is 0), with EQ the match word of the column's base.  They are not
// stored; the G lanes of a group re-derive them for G consecutive columns at once from the stored columns (one coalesced
// load per window instead of a chain of dependent loads per step), and the walk then runs on bit tests out of shared
// memory.  Everything else - first columns of nodes with several in-neighbours, slice borders, min-merged columns, the
// trimmed tail - takes the general path, which evaluates the reference's candidates from cell values.
//
// The walk itself is scalar per stream: all lanes of a group carry the same state, lane 0 writes the results.  The same
// source compiles for the host (oracle/hostsim): there the "lanes" of the window load are a loop.
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

// per group, in shared memory on the device
#define GA_TR_WIN 32        /* columns per window */
#define GA_TR_NODES 32      /* band nodes of the current slice kept in shared memory (larger bands are searched in global memory) */
template <int G>
struct GaTraceWindow
{
	uint64_t VP[GA_TR_WIN], VN[GA_TR_WIN];         // window load: the columns themselves (entry i = column top - i) ...
	uint64_t H[GA_TR_WIN], D0[GA_TR_WIN], EQ[GA_TR_WIN];   // ... and the masks of the word step that produced them
	uint32_t sw[GA_TR_WIN];                        // score word (score | flags); 0xffffffff = no such column
	uint32_t ok[GA_TR_WIN];                        // bit 0: H/D0 valid against the left neighbour in the node, bit 1: against the linked neighbour
	uint64_t peq[4];                               // match words of the current slice
	uint64_t nodeW[GA_TR_NODES];                   // current slice's band: first base of each node in the graph sequence,
	uint32_t nodeId[GA_TR_NODES], nodeCs[GA_TR_NODES], nodeLen[GA_TR_NODES];   // node, first column in the slab, length
};

#ifdef __CUDACC__
#define GA_TR_LANES(gl) if (const int gl = (int)(threadIdx.x & (G - 1)); true)
#define GA_TR_SYNC() __syncwarp(groupMask)
#define GA_TR_LEADER ((threadIdx.x & (G - 1)) == 0)
#define GA_TR_PREFETCH(p) asm volatile("prefetch.global.L2 [%0];" :: "l"(p))
#else
#define GA_TR_LANES(gl) for (int gl = 0; gl < G; gl++)
#define GA_TR_SYNC()
#define GA_TR_LEADER true
#define GA_TR_PREFETCH(p) (void)(p)
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
	for (uint32_t i = 0; i < nNodes; i++)
	{
		if (GA_TR_HN(nodeOff + i, 0) == node) return (int)i;
	}
	return -1;
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

// Walks one stream.  Called by all G lanes of the group with identical arguments; win is the group's window.
template <int G>
GA_DEV void ga_trace_stream(const ga_graph_view& g, const GaTraceMem& tm, GaTraceWindow<G>& win, uint32_t groupMask, const ga_stream_in& in, int nSlices, uint32_t node, uint32_t off,
	int32_t& statusOut, uint32_t& nMovesOut, uint32_t& nPathOut, uint32_t& nRunsOut, uint32_t& nPosOut)
{
	(void)groupMask;
	int32_t status = GA_OK;
	bool walking = true;
	uint32_t nMoves = 0, nPath = 0, nRuns = 0;
	// Same-node runs of the trimmed trace, walked backwards.  A run is opened at the first untrimmed position seen on a
	// node ('last' in read order) and closed when the walk leaves the node; its other end is the position left from.
	bool runOpen = false;
	uint32_t runNode = 0, runLastOff = 0, runLastRow = 0;
	uint32_t skipped = 0;        // positions dropped because their row lies in the trimmed tail
	uint32_t curWord = 0;
	int s = nSlices - 1;
	int row = 63;
	const int32_t maxv = (int32_t)in.partLen;
	const uint32_t startNode = in.startNode;
	const uint32_t trimRows = in.trimRows;
	// the slice the walk is in: header and (bands of up to GA_TR_NODES nodes) its node list in shared memory
	int loadedSlice = -1;
	uint32_t sNodeOff = 0, sNodes = 0, sSlab = 0;
	bool sCached = false;
	uint32_t slot = 0;           // band slot of `node` in slice s
	uint32_t colBase = 0;        // index of the node's first column in the history pool
	uint64_t nodeW = 0;          // first base of the node in the graph sequence
	bool reload = true;          // slice or node changed: re-resolve slot and colBase
	int32_t here = 0;
	bool haveHere = false;
	int winTop = -1;             // offset (in the node) of the window's newest column; -1 = no window
	const uint32_t maxMoves = tm.maxMoves;
	const bool leader = GA_TR_LEADER;
#define GA_TR_EMIT(mv) \
	{ \
		curWord |= (uint32_t)(mv) << ((nMoves & 15) * 2); \
		nMoves++; \
		if ((nMoves & 15) == 0) { if (leader) tm.moves[(size_t)((nMoves >> 4) - 1) * tm.S] = curWord; curWord = 0; } \
	}
	while (walking)
	{
		if (reload)
		{
			if (loadedSlice != s)
			{
				// ---- entering a slice: its header, match words and node list; the next slice's are requested into L2 ----
				sSlab = GA_TR_HDR(s, 0);
				sNodeOff = GA_TR_HDR(s, 2);
				sNodes = GA_TR_HDR(s, 3);
				sCached = sNodes <= GA_TR_NODES;
				GA_TR_SYNC();   // everybody is done with the old slice's tables
				GA_TR_LANES(gl)
				{
					if (gl < 2)
					{
						const uint4 q = tm.peq[(size_t)s * 2 + gl];
						win.peq[gl * 2] = (uint64_t)q.x | ((uint64_t)q.y << 32);
						win.peq[gl * 2 + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
					}
					if (sCached)
					{
						for (uint32_t i = (uint32_t)gl; i < sNodes; i += G)
						{
							const uint32_t nd = GA_TR_HN(sNodeOff + i, 0);
							win.nodeId[i] = nd;
							win.nodeCs[i] = GA_TR_HN(sNodeOff + i, 1);
							win.nodeLen[i] = GA_TR_HN(sNodeOff + i, 3);
							win.nodeW[i] = g.nodeStart[nd];
						}
					}
					if (s > 0 && gl == 0)
					{
						GA_TR_PREFETCH(&GA_TR_HDR(s - 1, 0));
						GA_TR_PREFETCH(tm.peq + (size_t)(s - 1) * 2);
					}
				}
				GA_TR_SYNC();
				loadedSlice = s;
			}
			int found = -1;
			if (sCached)
			{
				for (uint32_t i = 0; i < sNodes; i++) if (win.nodeId[i] == node) { found = (int)i; break; }
			}
			else found = ga_tr_find(tm, sNodeOff, sNodes, node);
			if (found < 0) { status = GA_ERR_TRACE; break; }
			slot = (uint32_t)found;
			if (sCached) { colBase = sSlab + win.nodeCs[slot]; nodeW = win.nodeW[slot]; }
			else { colBase = sSlab + GA_TR_HN(sNodeOff + slot, 1); nodeW = g.nodeStart[node]; }
			if (!haveHere)
			{
				const GaTrCol c = ga_tr_col(tm, colBase + off);
				here = ga_col_value(c.VP, c.VN, c.sbs, row);
				haveHere = true;
			}
			winTop = -1;
			reload = false;
		}
		if (!runOpen)
		{
			// still inside the trimmed tail, or a run was just closed: open one at the first untrimmed position
			const uint32_t j = (uint32_t)s * 64u + (uint32_t)row;
			if (j < trimRows) { runOpen = true; runNode = node; runLastOff = off; runLastRow = j; }
			else skipped++;
		}
		uint32_t okBits = 0;
		int wi = 0;
		if (runOpen && row > 0 && nMoves < maxMoves)
		{
			// ---- window: the GA_TR_WIN columns ending at `off`, their masks re-derived by the lanes in parallel ----------
			if (winTop < 0 || (int)off > winTop || winTop - (int)off >= GA_TR_WIN)
			{
				GA_TR_SYNC();   // everybody is done reading the old window
				winTop = (int)off;
				GA_TR_LANES(gl)
				{
#pragma unroll
					for (int r = 0; r < GA_TR_WIN / G; r++)
					{
						const int i = r * G + gl;
						const int c = winTop - i;
						uint32_t sw = 0xffffffffu;
						uint64_t vp = 0, vn = 0;
						// column -1 = the column stored right before the node's first one (GA_CF_LINK says when that means something)
						if (c >= 0 || (c == -1 && colBase > 0))
						{
							const uint32_t idx = (uint32_t)((int)colBase + c);
							const uint4 a = tm.colVV[(size_t)idx * tm.S];
							vp = (uint64_t)a.x | ((uint64_t)a.y << 32);
							vn = (uint64_t)a.z | ((uint64_t)a.w << 32);
							sw = tm.colS[(size_t)idx * tm.S];
						}
						win.VP[i] = vp;
						win.VN[i] = vn;
						win.sw[i] = sw;
					}
				}
				GA_TR_SYNC();
				GA_TR_LANES(gl)
				{
#pragma unroll
					for (int r = 0; r < GA_TR_WIN / G; r++)
					{
						const int i = r * G + gl;
						const int c = winTop - i;
						uint32_t ok = 0;
						uint64_t H = 0, D0 = 0, EQ = 0;
						const uint32_t sw = win.sw[i];
						if (c >= 0 && sw != 0xffffffffu && (sw & (c > 0 ? GA_CF_PLAIN : GA_CF_LINK)))
						{
							// the left neighbour: the next entry of the window, or one more load for the last entry
							uint64_t lvp, lvn;
							bool haveLeft = true;
							if (i + 1 < GA_TR_WIN)
							{
								lvp = win.VP[i + 1];
								lvn = win.VN[i + 1];
								haveLeft = win.sw[i + 1] != 0xffffffffu;
							}
							else
							{
								const uint32_t idx = (uint32_t)((int)colBase + c - 1);
								const uint4 a = tm.colVV[(size_t)idx * tm.S];
								lvp = (uint64_t)a.x | ((uint64_t)a.y << 32);
								lvn = (uint64_t)a.z | ((uint64_t)a.w << 32);
							}
							if (haveLeft)
							{
								const uint64_t w = nodeW + (uint64_t)c;
								const uint32_t base = (g.seq2[w >> 4] >> ((uint32_t)(w & 15) * 2)) & 3u;
								EQ = win.peq[base];
								// the horizontal half of ga_next_col with the match bit of row 0 as the forward pass used it
								const uint64_t Eq = (EQ & ~(uint64_t)1) | ((sw & GA_CF_EQ0) ? 1u : 0u);
								const uint64_t Xh = (((Eq & lvp) + lvp) ^ lvp) | Eq;
								H = lvn | ~(Xh | lvp);
								D0 = Xh | lvn;
								ok = c > 0 ? 1u : 2u;
							}
						}
						win.H[i] = H;
						win.D0[i] = D0;
						win.EQ[i] = EQ;
						win.ok[i] = ok;
					}
				}
				GA_TR_SYNC();
			}
			wi = winTop - (int)off;
			okBits = win.ok[wi];
		}
		// ---- fast step: inside the node, inside the slice, on a plain word-step column (nine steps in ten).  The three
		// candidates of pickBacktracePredecessor reduce to three bit tests, taken in the reference's order: horizontal,
		// diagonal, vertical.  Nothing here can fail or change node / slice / run.
		if ((okBits & 1u) && off > 0)
		{
			const uint32_t hbit = (uint32_t)(win.H[wi] >> row) & 1u;
			const uint32_t d0 = (uint32_t)(win.D0[wi] >> row) & 1u;
			const uint32_t eq = (uint32_t)(win.EQ[wi] >> row) & 1u;
			// diagonal delta = 1 - D0 must equal the mismatch cost 1 - EQ
			const uint32_t mv = hbit ? (uint32_t)GA_MOVE_H : (d0 == eq ? (uint32_t)GA_MOVE_D : (uint32_t)GA_MOVE_V);
			here = here - 1 + (int32_t)((mv == GA_MOVE_D) ? d0 : 0u);
			GA_TR_EMIT(mv);
			if (mv != GA_MOVE_V) off--;
			row -= (mv != GA_MOVE_H) ? 1 : 0;
			continue;
		}
		// ---- link step: first column of a node whose only band in-neighbour is in this slice and stored right before it.
		// The column is a word step from that neighbour's last column, so the same three bit tests apply (in-neighbour
		// horizontal, in-neighbour diagonal, vertical: GraphAligner.h:501-533 with one neighbour).
		// here < maxv: an in-neighbour outside the band reads as maxv in the reference and must not be able to match.
		if ((okBits & 2u) && off == 0 && here < maxv && nPath < tm.maxPathNodes && nRuns < tm.maxRuns && slot > 0)
		{
			const uint32_t hbit = (uint32_t)(win.H[wi] >> row) & 1u;
			const uint32_t d0 = (uint32_t)(win.D0[wi] >> row) & 1u;
			const uint32_t eq = (uint32_t)(win.EQ[wi] >> row) & 1u;
			const uint32_t mv = hbit ? (uint32_t)GA_MOVE_H : (d0 == eq ? (uint32_t)GA_MOVE_D : (uint32_t)GA_MOVE_V);
			here = here - 1 + (int32_t)((mv == GA_MOVE_D) ? d0 : 0u);
			GA_TR_EMIT(mv);
			if (mv != GA_MOVE_V)
			{
				// leaving the node: close the run on the position we stand on, cross into the neighbour (the band slot before)
				if (leader)
				{
					uint32_t* r = tm.runs + (size_t)(nRuns * GA_RUN_WORDS) * tm.S;
					r[0] = runNode; r[tm.S] = 0; r[2 * tm.S] = runLastOff; r[3 * tm.S] = (uint32_t)s * 64u + (uint32_t)row; r[4 * tm.S] = runLastRow;
				}
				nRuns++;
				runOpen = false;
				slot--;
				uint32_t len;
				if (sCached) { node = win.nodeId[slot]; len = win.nodeLen[slot]; nodeW = win.nodeW[slot]; }
				else { node = GA_TR_HN(sNodeOff + slot, 0); len = GA_TR_HN(sNodeOff + slot, 3); nodeW = g.nodeStart[node]; }
				off = len - 1;
				colBase -= len;
				if (leader) tm.pathNodes[(size_t)nPath * tm.S] = node;
				nPath++;
				winTop = -1;
			}
			row -= (mv != GA_MOVE_H) ? 1 : 0;
			continue;
		}
		// ---- general path (node starts, slice borders, merged columns, trimmed tail): the reference's candidates from
		// the stored columns' cell values
		uint32_t move = 4;
		uint32_t nnode = node, noff = off;
		int32_t nhere = 0;
		{
			const GaTrCol cur = ga_tr_col(tm, colBase + off);
			const uint64_t w = nodeW + off;
			const uint32_t base = ga_base(g, w);
			const int32_t match = (int32_t)((win.peq[base] >> row) & 1);
			const int32_t diagWant = here - 1 + match;
			const bool firstRow = s == 0 && row == 0;
			if (firstRow && node == startNode && (here == 0 || here == 1))
			{
				move = GA_MOVE_END;   // GraphAligner.h:500
			}
			else if (off > 0)
			{
				const GaTrCol left = ga_tr_col(tm, colBase + off - 1);
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
					const uint32_t upOff = GA_TR_HDR(s - 1, 2), upNodes = GA_TR_HDR(s - 1, 3);
					const int upSlot = ga_tr_find(tm, upOff, upNodes, node);
					if (upSlot < 0) ds = us = maxv;
					else
					{
						const uint32_t upBase = GA_TR_HDR(s - 1, 0) + GA_TR_HN(upOff + upSlot, 1);
						ds = ga_tr_end_score(tm, upBase + off - 1);
						us = ga_tr_end_score(tm, upBase + off);
					}
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
					if (sCached)
					{
						for (uint32_t i = 0; i < sNodes; i++) if (win.nodeId[i] == u) { uslot = (int)i; break; }
						uoff = uslot >= 0 ? win.nodeLen[uslot] - 1 : (uint32_t)(g.nodeStart[u + 1] - g.nodeStart[u]) - 1;
					}
					else
					{
						uslot = ga_tr_find(tm, sNodeOff, sNodes, u);
						uoff = (uint32_t)(g.nodeStart[u + 1] - g.nodeStart[u]) - 1;
					}
					GaTrCol uc;
					uc.VP = uc.VN = 0; uc.sbs = 0;
					if (uslot >= 0) uc = ga_tr_col(tm, sSlab + (sCached ? win.nodeCs[uslot] : GA_TR_HN(sNodeOff + uslot, 1)) + uoff);
					const int32_t hs = uslot >= 0 ? ga_col_value(uc.VP, uc.VN, uc.sbs, row) : maxv;
					if (hs == here - 1) { move = GA_MOVE_H; nnode = u; noff = uoff; nhere = hs; break; }
					int32_t ds;
					if (row == 0) ds = ga_tr_value(tm, startNode, s - 1, u, uoff, 63, maxv);
					else ds = uslot >= 0 ? ga_col_value(uc.VP, uc.VN, uc.sbs, row - 1) : maxv;
					if (ds == diagWant) { move = GA_MOVE_D; nnode = u; noff = uoff; nhere = ds; break; }
				}
				if (move == 4)
				{
					int32_t us;
					if (row > 0) us = here - (int32_t)((cur.VP >> row) & 1) + (int32_t)((cur.VN >> row) & 1);
					else us = ga_tr_value(tm, startNode, s - 1, node, off, 63, maxv);
					if (us == here - 1) { move = GA_MOVE_V; nhere = us; }
				}
			}
			// any step into row -1 ends the trace; that last position is popped again (GraphAligner.h:949-951)
			if (firstRow && (move == GA_MOVE_D || move == GA_MOVE_V)) move = GA_MOVE_END;
			if (move == 4) { status = GA_ERR_TRACE; break; }   // reference: assert(false); std::abort()
		}
		if (nMoves >= maxMoves) { status = GA_ERR_TRACE_OVERFLOW; break; }
		GA_TR_EMIT(move);
		// leaving the node (or ending): close the open run; its first position is the one we stand on
		if (runOpen && (move == GA_MOVE_END || nnode != node))
		{
			if (nRuns >= tm.maxRuns) { status = GA_ERR_TRACE_OVERFLOW; break; }
			if (leader)
			{
				uint32_t* r = tm.runs + (size_t)(nRuns * GA_RUN_WORDS) * tm.S;
				r[0] = runNode; r[tm.S] = off; r[2 * tm.S] = runLastOff; r[3 * tm.S] = (uint32_t)s * 64u + (uint32_t)row; r[4 * tm.S] = runLastRow;
			}
			nRuns++;
			runOpen = false;
		}
		if (move == GA_MOVE_END) { walking = false; continue; }
		here = nhere;
		if (move != GA_MOVE_V && off == 0)
		{
			if (nPath >= tm.maxPathNodes) { status = GA_ERR_TRACE_OVERFLOW; break; }
			if (leader) tm.pathNodes[(size_t)nPath * tm.S] = nnode;
			nPath++;
			reload = true;
		}
		if (move != GA_MOVE_H)
		{
			row--;
			if (row < 0)
			{
				row = 63;
				s--;
				reload = true;
			}
		}
		node = nnode;
		off = noff;
	}
#undef GA_TR_EMIT
	if (leader && (nMoves & 15)) tm.moves[(size_t)(nMoves >> 4) * tm.S] = curWord;
	statusOut = status;
	nMovesOut = nMoves;
	nPathOut = nPath;
	nRunsOut = nRuns;
	// positions = the start cell plus one per move except the terminating one, minus the trimmed tail
	nPosOut = nMoves > skipped ? nMoves - skipped : 0;
}

#endif
