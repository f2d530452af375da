// Per-stream banded bit-parallel DP + traceback: the device-side algorithm.
//
// One GPU thread owns one DP stream (one direction of one read/seed pair); the LANES (<= 32) streams of a warp
// run the slice loop in lock step so that their column history can be interleaved lane-by-lane (coalesced
// 16-byte stores) and so that __reduce_*_sync can size the shared slab.  LANES is chosen per batch: 32 when
// there are enough streams to fill the GPU, fewer when the batch is small and per-stream latency decides.  All arithmetic is integer/bitwise except the
// 2-state correctness HMM, which only adds/compares host-precomputed doubles (no FMA contraction possible).
//
// What is reproduced (reference = /root/reference):
//   band selection            projectForwardFromMinScore                    GraphAligner.h:1110-1159
//   row -1 scores + flags     forceComponentZeroRow                         GraphAligner.h:1903-1995
//   word step                 getNextSlice (Myers VP/VN)                    GraphAligner.h:1349-1427
//   node fill / in-edge merge calculateNode, getNodeStartSlice, sources     GraphAligner.h:1270-1347,1457-1573
//   column min-merge          WordSlice::mergeTwoSlices (values only)       WordSlice.h:361-421
//   slice driver / stop rule  getSqrtSlices, removeWronglyAlignedEnd        GraphAligner.h:2554-2856
//   correctness HMM           AlignmentCorrectnessEstimationState::NextState AlignmentCorrectnessEstimation.cpp:71-89
//   traceback + tie-breaks    getTraceFromTable*, pickBacktracePredecessor  GraphAligner.h:493-591,894-1021
// What is deliberately different: no confirmedRows bookkeeping on acyclic band components (final after one
// topological pass; cyclic components replay the reference's work list), no sqrt checkpointing (the column history
// - VP, VN and the row -1 score, 20 bytes per column - stays in HBM, so the traceback never recomputes a slice),
// no graph-sized scratch.  The traceback is a kernel of its own (ga_trace.cuh): this file is the forward pass.
#ifndef GA_CORE_CUH
#define GA_CORE_CUH
#include <stdint.h>
#include <string.h>
#include "ga_types.h"

#ifdef __CUDACC__
#define GA_DEV __device__ __forceinline__
#define GA_DEV_NOINLINE __device__ __forceinline__   /* a real call would cap the kernel at 96 registers (ABI) and spill in the hot loops */
#define GA_POPC(x) __popcll(x)
#define GA_CTZ(x) (__ffsll((long long)(x)) - 1)
#define GA_WARP_MAX(x) __reduce_max_sync(0xffffffffu, (x))
#define GA_WARP_ANY(x) __any_sync(0xffffffffu, (x))
#define GA_SYNCWARP() __syncwarp()
// one allocation per warp from a global bump pointer; every lane gets the same offset
#define GA_POOL_ALLOC(ptr, n) __shfl_sync(0xffffffffu, ((threadIdx.x & 31) == 0) ? atomicAdd((ptr), (unsigned long long)(n)) : 0ull, 0)
#else
#define GA_DEV inline
#define GA_DEV_NOINLINE inline
#define GA_POPC(x) __builtin_popcountll(x)
#define GA_CTZ(x) __builtin_ctzll(x)
#define GA_WARP_MAX(x) (x)
#define GA_WARP_ANY(x) (x)
#define GA_SYNCWARP()
#define GA_POOL_ALLOC(ptr, n) ((*(ptr) += (n)) - (n))
#endif

// -DGA_PHASE_TIMING: per-stream cycle counters per phase (profiling builds only: GA_EXTRA_FLAGS in build.py; the sums are
// printed by FinishStaged).  A lane's counter also collects the time it waits for the other lanes of its warp.
#if defined(GA_PHASE_TIMING) && defined(__CUDACC__)
#define GA_T0(st) (st).tLast = clock64()
#define GA_TLAP(st, i) { long long n_ = clock64(); (st).phase[i] += (unsigned long long)(n_ - (st).tLast); (st).tLast = n_; }
#else
#define GA_T0(st)
#define GA_TLAP(st, i)
#endif

#define GA_ALT_CUTOFF 200000u   // GraphAlignerCommon.h:10
#define GA_HDR_WORDS 12u        // slabOff, ncols, nodeOff, nNodes, minScore, flags, HMM state after the slice (2 doubles), last minimum cell (slot, column)
#define GA_HN_WORDS 5u          // node, colStart, nodeMin, len, first sequence chunk (ga_node_rec::seqChunk)
// flags word of a slice header (word 5): bit 0 CurrentlyCorrect, bit 1 FalseFromCorrect, and for -B ramp runs (ga_run_stream<.., RAMP>):
#define GA_HF_ALT 4u            // the reference kept a sqrt checkpoint of this slice that is another instance of it than the one the
                                // history holds: word 1 = header index of that instance (what the slice looks like from the slice below)
#define GA_HF_RAMPBW 8u         // the slice ran with rampBandwidth (DPTable::bandwidthPerSlice)
#define GA_HF_STALE 16u         // (checkpoint entries) the entry is not the instance the forward pass ended with
// -B ramp runs keep the reference's sqrt checkpoints (DPTable::slices, GraphAligner.h:2772-2786) as copies of slice headers behind the
// maxSlices headers of a stream: GA_CP_SLOTS stack entries, then the pending one (storeSlice); word 1 of an entry = slice index + 1
#define GA_CP_EXTRA 8u
#define GA_CP_SLOTS(maxSlices) ((maxSlices) + GA_CP_EXTRA)
#define GA_HDR_SLOTS_RAMP(maxSlices) (2u * (maxSlices) + GA_CP_EXTRA + 1u)

#ifdef GA_HOST_DEBUG
static unsigned long long g_dbgFast = 0, g_dbgOuter = 0, g_dbgGeneral = 0, g_dbgNodeStart = 0, g_dbgRow0 = 0, g_dbgMerged = 0, g_dbgReload = 0, g_dbgLink = 0;
#endif

struct GaHmmTables
{
	double correctMul[65];
	double falseMul[65];
	double c2c, c2f, f2c, f2f;
	double startCorrect, startFalse;
};

#if !defined(__CUDACC__) && !defined(GA_HOSTSIM)
struct uint4 { uint32_t x, y, z, w; };
#endif

// per-lane memory; every pointer is already offset by the lane, element i lives at p[i * LANES]
struct GaLaneMem
{
	void* tiny[2];       // frozen end state per band column, current / previous slice (ga_tiny_ld / ga_tiny_st)
	uint64_t* hash[2];   // node -> band slot, (node << 32 | stamp << 16 | slot)
	uint64_t* heap;
	uint32_t* indeg;
	uint32_t* order;
	uint32_t* nWlo;      // per band node of the current slice: nodeStart (low / high word) and its first column
	uint32_t* nWhi;      //   in the previous slice's tiny array (0xffffffff = not in the previous band)
	uint32_t* nPcs;
	uint32_t* cmpOf;     // cyclic slices only: Tarjan component of each band node (top bit: on the work list)
	uint32_t* emit;      //   band slots in Tarjan emission order (components are contiguous)
	uint32_t* wl;        //   the reference's UniqueQueue (a LIFO)
	uint32_t* conf;      //   per band column: confirmedRows (rows | partial << 8), GraphAligner.h:1355-1416
	uint32_t* hdr;
	uint32_t* histNode;
	uint32_t* hnRing;    // small-band mode: node lists of the current and the previous slice (GA_HN)
	uint64_t* lastVV;    // small-band mode: {VP, VN} and the row -1 score of the last column of every band node evaluated
	uint32_t* lastS;     //   in this slice (what the first column of a successor starts from)
	uint64_t* eqTab;     // four match words of the current slice, [base][lane] (shared memory on the device)
	uint4* colVV;        // column history pool (shared by all warps): {VP, VN} per column, [column][lane]
	uint32_t* colS;      //   and the row -1 score with the traceback's flags (GA_CF_*), [column][lane]
	unsigned long long* colPoolTop;   // bump pointer of the pool, in columns (x LANES lanes)
	uint32_t* ubkt;      // unordered_map emulation: bucket -> "before" node
	uint32_t* unext;     // unordered_map emulation: forward list links
	uint32_t* uorder;    // iteration order of the previous slice's node map
	const uint4* peq;    // this stream's match masks, two 16-byte halves per slice: {A, C} and {G, T} (not interleaved)
};

#define GA_HDR(s, f) mem.hdr[(size_t)((s) * GA_HDR_WORDS + (f)) * LANES]
// Node lists of the slices (node history).  GA_HNG = the history in global memory.  GA_HN = the same entry as the forward
// pass sees it: in small-band mode the lists of the current and the previous slice (adjacent in the history, at most
// GA_SMEM_NODES entries each) live in a shared-memory ring of GA_HN_RING entries indexed by the entry number, and a slice's
// list is copied to the history when the slice is kept (ga_run_stream).
#define GA_HN_RING 32u
#define GA_HNG(i, f) mem.histNode[(size_t)((i) * GA_HN_WORDS + (f)) * LANES]
#define GA_HN(i, f) (*(SMALL ? &mem.hnRing[(size_t)((((i) & (GA_HN_RING - 1u)) * GA_HN_WORDS) + (f)) * LANES] : &GA_HNG(i, f)))
#define GA_NOT_IN_PREV 0xffffffffu

struct GaCol
{
	uint64_t VP, VN;
	int32_t sbs, scoreEnd;
};

// Column history record: {VP, VN} (16 bytes) + one word = row -1 score (29 bits) | flags.  The end score is not stored:
// scoreEnd = sbs + popcount(VP) - popcount(VN) (reference invariant, GraphAligner.h:1421).
//   GA_CF_PLAIN  the column is a plain word step (ga_next_col, no min-merge) from its left neighbour in the node, so the
//                traceback may decide its steps from the step's own bit masks (ga_trace.cuh) instead of cell values
//   GA_CF_LINK   first column of a node: a plain word step from the last column of the node's ONLY band in-neighbour, and
//                that column is the one stored right before this one in the slice's slab
//   GA_CF_EQ0    bit 0 of the match word as the step used it (the band-edge rules and a negative horizontal entry change it)
#define GA_CF_PLAIN 0x80000000u
#define GA_CF_EQ0 0x40000000u
#define GA_CF_LINK 0x20000000u
#define GA_CF_SCORE_MASK 0x1fffffffu

template <int LANES>
GA_DEV void ga_col_store(const GaLaneMem& mem, uint32_t col, const GaCol& c, uint32_t flags)
{
	uint4 a;
	a.x = (uint32_t)c.VP; a.y = (uint32_t)(c.VP >> 32); a.z = (uint32_t)c.VN; a.w = (uint32_t)(c.VN >> 32);
	mem.colVV[(size_t)col * LANES] = a;
	mem.colS[(size_t)col * LANES] = (uint32_t)c.sbs | flags;
}

template <int LANES>
GA_DEV GaCol ga_col_load(const GaLaneMem& mem, uint32_t col)
{
	uint4 a = mem.colVV[(size_t)col * LANES];
	GaCol c;
	c.VP = (uint64_t)a.x | ((uint64_t)a.y << 32);
	c.VN = (uint64_t)a.z | ((uint64_t)a.w << 32);
	c.sbs = (int32_t)(mem.colS[(size_t)col * LANES] & GA_CF_SCORE_MASK);
	c.scoreEnd = c.sbs + (int32_t)GA_POPC(c.VP) - (int32_t)GA_POPC(c.VN);
	return c;
}

template <int LANES>
GA_DEV int32_t ga_col_load_sbs(const GaLaneMem& mem, uint32_t col)
{
	return (int32_t)(mem.colS[(size_t)col * LANES] & GA_CF_SCORE_MASK);
}

template <int LANES>
GA_DEV void ga_col_store_sbs(const GaLaneMem& mem, uint32_t col, int32_t sbs)
{
	mem.colS[(size_t)col * LANES] = (uint32_t)sbs;
}

// tiny = frozen end state of a column, cf. reference TinySlice (NodeSlice.h:26-31):
// bit0 VP63, bit1 VN63, bit2 scoreBeforeExists of the column, bits 3.. scoreEnd
GA_DEV uint32_t ga_tiny_pack(const GaCol& c, bool sbE)
{
	return ((uint32_t)c.scoreEnd << 3) | (sbE ? 4u : 0u) | (uint32_t)((c.VN >> 62) & 2) | (uint32_t)(c.VP >> 63);
}
GA_DEV int32_t ga_tiny_score(uint32_t t) { return (int32_t)(t >> 3); }
// value of row 62 of the frozen column = scoreEnd - VP63 + VN63 (GraphAligner.h:1368)
GA_DEV int32_t ga_tiny_row62(uint32_t t) { return (int32_t)(t >> 3) - (int32_t)(t & 1) + (int32_t)((t >> 1) & 1); }

// The tiny arrays are 32-bit words in global memory in the general layout.  In small-band mode (SMALL) they live in shared
// memory as 16-bit words: the three flags and the low 13 bits of the score, which is rebuilt against a reference score that
// is known to be at most GA_TINY_SPAN below it (the minimum of the previous slice: no end score of this or the previous
// slice is lower).  A stream whose scores spread further leaves small-band mode (ga_node_columns).
#define GA_TINY_SPAN 8000
template <int LANES, bool SMALL>
GA_DEV uint32_t ga_tiny_ld(const void* base, uint32_t idx, int32_t ref)
{
	if (SMALL)
	{
		const uint32_t t = ((const uint16_t*)base)[(size_t)idx * LANES];
		const uint32_t score = (uint32_t)ref + (((t >> 3) - (uint32_t)ref) & 0x1fffu);
		return (score << 3) | (t & 7u);
	}
	return ((const uint32_t*)base)[(size_t)idx * LANES];
}
template <int LANES, bool SMALL>
GA_DEV void ga_tiny_st(void* base, uint32_t idx, uint32_t t)
{
	if (SMALL) ((uint16_t*)base)[(size_t)idx * LANES] = (uint16_t)t;
	else ((uint32_t*)base)[(size_t)idx * LANES] = t;
}
#define GA_TP(x) ga_tiny_ld<LANES, SMALL>(cx.tinyPrev, (x), cx.tinyRef)
#define GA_TC(x) ga_tiny_ld<LANES, SMALL>(cx.tinyCur, (x), cx.tinyRef)
#define GA_TC_ST(x, v) ga_tiny_st<LANES, SMALL>(cx.tinyCur, (x), (v))

GA_DEV uint32_t ga_base(const ga_graph_view& g, uint64_t w)
{
	return (g.seq2[w >> 4] >> ((uint32_t)(w & 15) * 2)) & 3u;
}

// Myers word step without the confirmedRows bookkeeping (GraphAligner.h:1349-1399).
// topScore = end score of this column in the previous slice (or INT_MAX): the reference takes the word step from
// the left neighbour and then min-merges the result with the vertical ramp from topScore when the step's row -1 score
// is larger (GraphAligner.h:1541-1546).  That minimum IS the word step whose row -1 score is topScore - same
// recurrence, lower entry value - so whenever topScore - L.sbs is a legal horizontal delta (-1 or 0) the step is taken
// with it directly, and the column stays a plain word step for the traceback.  needMerge reports the rare other case.
// eq0 = bit 0 of the match word as the horizontal part of the step used it (the traceback re-derives the step's masks).
GA_DEV GaCol ga_next_col(uint64_t Eq, const GaCol& L, bool leftSbE, bool upleftInside, bool diagInside, bool previousEq, int32_t upleftRow62, int32_t topScore, uint32_t& eq0, bool& needMerge)
{
	GaCol r;
	if (!leftSbE || !diagInside) Eq &= ~(uint64_t)1;
	int32_t sbs = L.sbs + 1;
	if (upleftInside)
	{
		int32_t d = upleftRow62 + (previousEq ? 0 : 1);
		if (d < sbs) sbs = d;
	}
	needMerge = false;
	if (topScore < sbs)
	{
		if (topScore - L.sbs >= -1) sbs = topScore;
		else needMerge = true;
	}
	int32_t hin = sbs - L.sbs;
	uint64_t Xv = Eq | L.VN;
	if (hin < 0) Eq |= 1;
	eq0 = (uint32_t)Eq & 1u;
	uint64_t Xh = (((Eq & L.VP) + L.VP) ^ L.VP) | Eq;
	uint64_t Ph = L.VN | ~(Xh | L.VP);
	uint64_t Mh = L.VP & Xh;
	r.scoreEnd = L.scoreEnd + (int32_t)(Ph >> 63) - (int32_t)(Mh >> 63);
	Ph <<= 1;
	Mh <<= 1;
	if (hin < 0) Mh |= 1; else if (hin > 0) Ph |= 1;
	r.VP = Mh | ~(Xv | Ph);
	r.VN = Ph & Xv;
	r.sbs = sbs;
	return r;
}

// column := min(column, vertical ramp from `top`) where top < column.sbs (GraphAligner.h:1504-1509,1541-1546
// with WordSlice.h:361-421).  The ramp is the steepest column there is, so the minimum differs from the
// computed column only by "paying back" D = sbs - top deficits at the first non-VP rows.
GA_DEV void ga_vertical_merge(GaCol& c, int32_t top)
{
	while (c.sbs > top)
	{
		uint64_t m = ~c.VP;
		if (m == 0)
		{
			c.scoreEnd -= 1;
		}
		else
		{
			uint64_t b = m & (0 - m);
			if (c.VN & b) c.VN ^= b; else c.VP |= b;
		}
		c.sbs -= 1;
	}
}

// exact element-wise minimum of two columns (values of WordSlice::mergeTwoSlices, WordSlice.h:361-421).
// A 1-Lipschitz column has a unique (sbs,VP,VN) form, so any exact minimum is bit-identical to the reference's.
// The difference d = A - B only changes at rows where the two columns' vertical deltas differ; everywhere else the
// minimum's delta is the common delta whichever column is lower.  So the result starts as A with those rows cleared, and
// only the differing rows (a handful for the columns that meet at a node start) are walked, d carried along.
GA_DEV_NOINLINE GaCol ga_merge_cols(const GaCol& A, const GaCol& B)
{
	GaCol r;
	r.sbs = A.sbs < B.sbs ? A.sbs : B.sbs;
	r.scoreEnd = A.scoreEnd < B.scoreEnd ? A.scoreEnd : B.scoreEnd;
	int32_t d = A.sbs - B.sbs;
	uint64_t diff = (A.VP ^ B.VP) | (A.VN ^ B.VN);
	uint64_t VP = A.VP & ~diff, VN = A.VN & ~diff;
	while (diff)
	{
		const uint64_t bit = diff & (0 - diff);
		diff ^= bit;
		const int32_t da = ((A.VP & bit) ? 1 : 0) - ((A.VN & bit) ? 1 : 0);
		const int32_t db = ((B.VP & bit) ? 1 : 0) - ((B.VN & bit) ? 1 : 0);
		const int32_t dn = d + da - db;
		int32_t delta;
		if (d <= 0) delta = (dn <= 0) ? da : da - dn;
		else delta = (dn > 0) ? db : db + dn;
		if (delta > 0) VP |= bit;
		else if (delta < 0) VN |= bit;
		d = dn;
	}
	r.VP = VP;
	r.VN = VN;
	return r;
}

GA_DEV uint64_t ga_double_to_bits(double d)
{
#ifdef __CUDACC__
	return (uint64_t)__double_as_longlong(d);
#else
	uint64_t u;
	memcpy(&u, &d, sizeof(u));
	return u;
#endif
}

GA_DEV double ga_bits_to_double(uint64_t u)
{
#ifdef __CUDACC__
	return __longlong_as_double((long long)u);
#else
	double d;
	memcpy(&d, &u, sizeof(d));
	return d;
#endif
}

GA_DEV int32_t ga_col_value(uint64_t VP, uint64_t VN, int32_t sbs, int row)
{
	uint64_t mask = (row >= 63) ? ~(uint64_t)0 : ~(~(uint64_t)0 << (row + 1));
	return sbs + (int32_t)GA_POPC(VP & mask) - (int32_t)GA_POPC(VN & mask);
}

// ---- open-addressing node -> band-slot table, stamped per slice so it never needs clearing ----------------
template <int LANES>
GA_DEV int ga_hash_find(const uint64_t* table, uint32_t hashMask, uint32_t stamp, uint32_t node)
{
	uint32_t h = ((node * 2654435761u) >> 15) & hashMask;
	while (true)
	{
		uint64_t e = table[(size_t)h * LANES];
		if (((uint32_t)(e >> 16) & 0xffffu) != stamp) return -1;
		if ((uint32_t)(e >> 32) == node) return (int)(e & 0xffffu);
		h = (h + 1) & hashMask;
	}
}

template <int LANES>
GA_DEV void ga_hash_insert(uint64_t* table, uint32_t hashMask, uint32_t stamp, uint32_t node, uint32_t slot)
{
	uint32_t h = ((node * 2654435761u) >> 15) & hashMask;
	while (((uint32_t)(table[(size_t)h * LANES] >> 16) & 0xffffu) == stamp) h = (h + 1) & hashMask;
	table[(size_t)h * LANES] = ((uint64_t)node << 32) | ((uint64_t)stamp << 16) | slot;
}

// A band is a handful of nodes, the table is sized for the largest band allowed: a slice only uses the first `window`
// entries (mask = window - 1), so the tables of all resident warps stay in L1.  Lookups for a stamp must use the mask the
// stamp was inserted with; ga_band_add widens the window (under a fresh stamp) when a band outgrows it.
GA_DEV uint32_t ga_hash_window(uint32_t expectedNodes, uint32_t hashSize)
{
	uint32_t size = 32;
	while (size < expectedNodes * 4 && size < hashSize) size <<= 1;
	if (size > hashSize) size = hashSize;
	return size - 1;
}

// ---- std::priority_queue<NodeWithPriority, vector, greater<>> as libstdc++ implements it ----------------------
// Entries are (priority << 32 | node) but ONLY the priority is compared (GraphAligner.h:1094-1108), and the
// sift order follows std::__push_heap / std::__adjust_heap step by step: the pop order among equal priorities
// decides the band's node order, which decides which of several tied minimum cells the traceback starts from.
template <int LANES>
GA_DEV void ga_heap_sift_up(uint64_t* heap, uint32_t hole, uint32_t top, uint64_t v)
{
	while (hole > top)
	{
		uint32_t parent = (hole - 1) >> 1;
		uint64_t pv = heap[(size_t)parent * LANES];
		if (!((uint32_t)(pv >> 32) > (uint32_t)(v >> 32))) break;
		heap[(size_t)hole * LANES] = pv;
		hole = parent;
	}
	heap[(size_t)hole * LANES] = v;
}

template <int LANES>
GA_DEV void ga_heap_push(uint64_t* heap, uint32_t& n, uint64_t v)
{
	ga_heap_sift_up<LANES>(heap, n, 0, v);
	n++;
}

template <int LANES>
GA_DEV uint64_t ga_heap_pop(uint64_t* heap, uint32_t& n)
{
	uint64_t top = heap[0];
	uint32_t len = --n;            // elements that stay in the heap
	if (len == 0) return top;
	uint64_t v = heap[(size_t)len * LANES];
	uint32_t hole = 0;
	uint32_t child = 0;
	while (child < (len - 1) / 2)
	{
		child = 2 * (child + 1);
		uint64_t r = heap[(size_t)child * LANES];
		uint64_t l = heap[(size_t)(child - 1) * LANES];
		if ((uint32_t)(r >> 32) > (uint32_t)(l >> 32)) { child--; r = l; }
		heap[(size_t)hole * LANES] = r;
		hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2)
	{
		child = 2 * (child + 1);
		heap[(size_t)hole * LANES] = heap[(size_t)(child - 1) * LANES];
		hole = child - 1;
	}
	ga_heap_sift_up<LANES>(heap, hole, 0, v);
	return top;
}

// ---- iteration order of a libstdc++ std::unordered_map<size_t,...> filled key by key ---------------------------
// The reference walks the previous slice's node map (NodeSlice.h:730-733 fills it in band order,
// GraphAligner.h:1117 iterates it); the walk order is a pure function of the insertion order and of the
// library's bucket-count schedule, which the host probes from a real std::unordered_map at start-up.
struct GaUmapSchedule
{
	uint32_t n;
	uint32_t threshold[48];   // inserting the threshold[i]-th element (1-based) rehashes to buckets[i] first
	uint32_t buckets[48];
};

#define GA_UB_EMPTY 0xffffffffu
#define GA_UB_BEGIN 0xfffffffeu
#define GA_UNIL 0xffffffffu

// keys: GA_HN(keyOff + i, 0), i in [0,n).  Writes the element indices in iteration order to mem.uorder.
template <int LANES, bool SMALL>
GA_DEV void ga_umap_order(const GaUmapSchedule& sch, const GaLaneMem& mem, uint32_t keyOff, uint32_t n)
{
	uint32_t bktCount = 1;
	uint32_t head = GA_UNIL;
	uint32_t si = 0;
	mem.ubkt[0] = GA_UB_EMPTY;
	for (uint32_t i = 0; i < n; i++)
	{
		if (si < sch.n && i + 1 == sch.threshold[si])
		{
			// _M_rehash_aux(n, true_type): relink every node, walking the old list front to back
			uint32_t nb = sch.buckets[si++];
			for (uint32_t b = 0; b < nb; b++) mem.ubkt[(size_t)b * LANES] = GA_UB_EMPTY;
			uint32_t p = head;
			head = GA_UNIL;
			uint32_t bbeginBkt = 0;
			while (p != GA_UNIL)
			{
				uint32_t nxt = mem.unext[(size_t)p * LANES];
				uint32_t b = GA_HN(keyOff + p, 0) % nb;
				uint32_t before = mem.ubkt[(size_t)b * LANES];
				if (before == GA_UB_EMPTY)
				{
					mem.unext[(size_t)p * LANES] = head;
					bool hadNext = head != GA_UNIL;
					head = p;
					mem.ubkt[(size_t)b * LANES] = GA_UB_BEGIN;
					if (hadNext) mem.ubkt[(size_t)bbeginBkt * LANES] = p;
					bbeginBkt = b;
				}
				else
				{
					uint32_t after = before == GA_UB_BEGIN ? head : mem.unext[(size_t)before * LANES];
					mem.unext[(size_t)p * LANES] = after;
					if (before == GA_UB_BEGIN) head = p; else mem.unext[(size_t)before * LANES] = p;
				}
				p = nxt;
			}
			bktCount = nb;
		}
		// _M_insert_bucket_begin
		uint32_t key = GA_HN(keyOff + i, 0);
		uint32_t b = key % bktCount;
		uint32_t before = mem.ubkt[(size_t)b * LANES];
		if (before != GA_UB_EMPTY)
		{
			uint32_t after = before == GA_UB_BEGIN ? head : mem.unext[(size_t)before * LANES];
			mem.unext[(size_t)i * LANES] = after;
			if (before == GA_UB_BEGIN) head = i; else mem.unext[(size_t)before * LANES] = i;
		}
		else
		{
			mem.unext[(size_t)i * LANES] = head;
			if (head != GA_UNIL) mem.ubkt[(size_t)(GA_HN(keyOff + head, 0) % bktCount) * LANES] = i;
			head = i;
			mem.ubkt[(size_t)b * LANES] = GA_UB_BEGIN;
		}
	}
	uint32_t k = 0;
	for (uint32_t p = head; p != GA_UNIL; p = mem.unext[(size_t)p * LANES]) mem.uorder[(size_t)(k++) * LANES] = p;
}

struct GaStreamState
{
	// stream constants
	const uint32_t* aux;   // per slice: exact code of the read character above the slice | IUPAC mask of the part's first character << 4 (ga_peq_kernel)
	uint32_t partLen;
	uint32_t nslices;
	uint32_t startNode;
	uint32_t trimRows;
	// running state
	int32_t status;
	bool done;
	int32_t prevMin;
	double hmmC, hmmF;
	uint32_t histNodeTop;
	uint32_t slicesPushed;
	uint64_t wordColumns;
	uint32_t cyclicSlices;
	uint32_t rampRedos;
#ifdef GA_PHASE_TIMING
	long long tLast;
	unsigned long long phase[16];  // see the names in FinishStaged
#endif
};

// IUPAC match masks (bit0 A, bit1 C, bit2 G, bit3 T), GraphAligner.h:2039-2110; 0 = invalid character
GA_DEV uint32_t ga_iupac_mask(uint8_t c)
{
	switch (c)
	{
		case 'A': case 'a': return 1;
		case 'C': case 'c': return 2;
		case 'G': case 'g': return 4;
		case 'T': case 't': return 8;
		case 'N': case 'n': return 15;
		case 'R': case 'r': return 1 | 4;
		case 'Y': case 'y': return 2 | 8;
		case 'K': case 'k': return 4 | 8;
		case 'M': case 'm': return 2 | 1;
		case 'S': case 's': return 2 | 4;
		case 'W': case 'w': return 1 | 8;
		case 'B': case 'b': return 2 | 4 | 8;
		case 'D': case 'd': return 1 | 4 | 8;
		case 'H': case 'h': return 1 | 2 | 8;
		case 'V': case 'v': return 1 | 2 | 4;
		default: return 0;
	}
}

// exact (case-sensitive, non-IUPAC) comparison used for the row above the slice, GraphAligner.h:1503,1540
GA_DEV uint32_t ga_exact_code(uint8_t c)
{
	switch (c)
	{
		case 'A': return 0;
		case 'C': return 1;
		case 'G': return 2;
		case 'T': return 3;
		default: return 4;
	}
}


// ------------------------------------------------------------------------------------------------------------
// Band selection for slice s from slice s-1 (GraphAligner.h:1110-1159).  Appends the band's node list to the
// node history at nodeOff (in the reference's band order), fills hashCur and the per-node scratch records.
// Returns the number of band nodes.
// ------------------------------------------------------------------------------------------------------------
template <int LANES, bool SMALL>
GA_DEV bool ga_band_add(const ga_graph_view& g, const ga_caps& caps, const GaLaneMem& mem, GaStreamState& st, uint64_t* hashCur, uint32_t& maskCur, uint32_t& stampCur, uint32_t& gen,
	uint32_t nodeOff, uint32_t& nc, uint32_t& ncols, uint32_t node, uint64_t wStart, uint32_t len, uint32_t pcs)
{
	if (nc >= caps.maxNodes) { st.status = GA_ERR_NODE_OVERFLOW; return false; }
	if (nodeOff + nc >= caps.histNodes) { st.status = GA_ERR_HIST_OVERFLOW; return false; }
	if ((nc + 1) * 2 > maskCur + 1 && maskCur + 1 < caps.hashSize)
	{
		// the table outgrew the window it was given (ga_hash_window): re-insert under a fresh stamp with a wider mask
		if (gen >= 0xfffeu) { st.status = GA_ERR_HIST_OVERFLOW; return false; }
		uint32_t size = maskCur + 1;
		while (size < (nc + 1) * 4 && size < caps.hashSize) size <<= 1;
		maskCur = size - 1;
		stampCur = ++gen;
		for (uint32_t i = 0; i < nc; i++) ga_hash_insert<LANES>(hashCur, maskCur, stampCur, GA_HN(nodeOff + i, 0), i);
	}
	GA_HN(nodeOff + nc, 0) = node;
	GA_HN(nodeOff + nc, 1) = ncols;
	GA_HN(nodeOff + nc, 3) = len;
	GA_HN(nodeOff + nc, 4) = g.nodeRec[node].seqChunk;
	mem.nWlo[(size_t)nc * LANES] = (uint32_t)wStart;
	mem.nWhi[(size_t)nc * LANES] = (uint32_t)(wStart >> 32);
	mem.nPcs[(size_t)nc * LANES] = pcs;
	ga_hash_insert<LANES>(hashCur, maskCur, stampCur, node, nc);
	nc++;
	ncols += len;
	if (ncols >= GA_ALT_CUTOFF) { st.status = GA_ERR_ALT_METHOD; return false; }
	return true;
}

template <int LANES, bool SMALL>
GA_DEV int ga_select_band(const ga_graph_view& g, const ga_caps& caps, const GaUmapSchedule& sch, const GaLaneMem& mem, GaStreamState& st, int bandwidth,
	uint32_t pNodeOff, uint32_t pNodes, const void* tinyPrev, const uint64_t* hashPrev, uint32_t maskPrev, uint32_t stampPrev, uint64_t* hashCur, uint32_t& maskCur, uint32_t& stampCur, uint32_t& gen,
	uint32_t nodeOff, uint32_t& ncolsOut)
{
	const int32_t expand = bandwidth + 64;
	uint32_t nc = 0;
	uint32_t ncols = 0;
	uint32_t heapN = 0;
	// the reference walks the previous slice's unordered_map (GraphAligner.h:1117)
	GA_TLAP(st, 0);
	ga_umap_order<LANES, SMALL>(sch, mem, pNodeOff, pNodes);
	GA_TLAP(st, 8);
	for (uint32_t it = 0; it < pNodes; it++)
	{
		const uint32_t i = mem.uorder[(size_t)it * LANES];
		int32_t nodeMin = (int32_t)GA_HN(pNodeOff + i, 2);
		if (nodeMin > st.prevMin + bandwidth) continue;
		uint32_t node = GA_HN(pNodeOff + i, 0);
		uint32_t pcs = GA_HN(pNodeOff + i, 1);
		uint32_t len = GA_HN(pNodeOff + i, 3);
		if (!ga_band_add<LANES, SMALL>(g, caps, mem, st, hashCur, maskCur, stampCur, gen, nodeOff, nc, ncols, node, g.nodeStart[node], len, pcs)) return -1;
		int32_t endscore = ga_tiny_score(ga_tiny_ld<LANES, SMALL>(tinyPrev, pcs + len - 1, st.prevMin));
		if (endscore > st.prevMin + expand) continue;
		for (uint32_t e = g.outOff[node], eEnd = g.outOff[node + 1]; e < eEnd; e++)
		{
			if (heapN >= caps.maxQueue) { st.status = GA_ERR_QUEUE_OVERFLOW; return -1; }
			ga_heap_push<LANES>(mem.heap, heapN, ((uint64_t)(uint32_t)(endscore - st.prevMin + 1) << 32) | g.outAdj[e]);
		}
	}
	GA_TLAP(st, 9);
	while (heapN > 0)
	{
		uint64_t top = mem.heap[0];
		int32_t prio = (int32_t)(top >> 32);
		if (prio > expand) break;
		ga_heap_pop<LANES>(mem.heap, heapN);
		uint32_t node = (uint32_t)top;
		if (ga_hash_find<LANES>(hashCur, maskCur, stampCur, node) >= 0) continue;
		uint64_t wStart = g.nodeStart[node];
		uint32_t len = (uint32_t)(g.nodeStart[node + 1] - wStart);
		// not kept, but it may still sit in the previous band (its minimum was outside the bandwidth)
		int pslot = ga_hash_find<LANES>(hashPrev, maskPrev, stampPrev, node);
		uint32_t pcs = pslot >= 0 ? GA_HN(pNodeOff + pslot, 1) : GA_NOT_IN_PREV;
		if (!ga_band_add<LANES, SMALL>(g, caps, mem, st, hashCur, maskCur, stampCur, gen, nodeOff, nc, ncols, node, wStart, len, pcs)) return -1;
		for (uint32_t e = g.outOff[node], eEnd = g.outOff[node + 1]; e < eEnd; e++)
		{
			if (heapN >= caps.maxQueue) { st.status = GA_ERR_QUEUE_OVERFLOW; return -1; }
			ga_heap_push<LANES>(mem.heap, heapN, ((uint64_t)(uint32_t)(prio + (int32_t)len) << 32) | g.outAdj[e]);
		}
	}
	GA_TLAP(st, 10);
	if (ncols > caps.maxCols) { st.status = GA_ERR_COL_OVERFLOW; return -1; }
	ncolsOut = ncols;
	return (int)nc;
}

struct GaSliceCtx
{
	int s;
	uint32_t nodeOff, nNodes;     // this slice's node list in the node history
	uint32_t pNodeOff, pNodes;    // previous slice's
	uint32_t slabOff;             // this slice's first column in the warp slab
	uint32_t pSlabOff;            // the previous slice's (unused when !hasPrevSlab: slice 0 follows the initial slice)
	bool hasPrevSlab;
	void* tinyCur;
	const void* tinyPrev;
	int32_t tinyRef;              // reference score of the 16-bit tiny encoding (the previous slice's minimum)
	uint64_t* hashCur;
	const uint64_t* hashPrev;
	uint32_t stampCur, stampPrev;
	uint32_t maskCur, maskPrev;   // hash windows of the two tables (ga_hash_window)
	uint64_t BA, BC, BG, BT;
	uint64_t* eqTab;              // the same four words indexed by base: [base][lane], shared memory on the device
	uint32_t prevCharCode;        // exact code of sequence[j0-1], 4 = matches nothing
	bool firstSlice;
};

// Columns 1..len-1 of a node: the serial Myers chain (GraphAligner.h:1532-1570).  INPREV = the node is also in
// the previous slice's band (then every column may be min-merged with the vertical ramp from the previous slice).
// Returns the minimum scoreEnd over the node.
template <int LANES, bool SMALL, bool INPREV>
GA_DEV int32_t ga_node_columns(const ga_graph_view& g, const GaLaneMem& mem, const GaSliceCtx& cx, uint64_t wStart, uint32_t len, uint32_t cs, uint32_t pcs,
	uint32_t prevMask, GaCol& L, bool LsbE, uint32_t oldTinyLeft, uint32_t oldTinyNext, int32_t nodeMin, bool& spanOverflow)
{
	uint4* vvPtr = mem.colVV + (size_t)(cx.slabOff + cs + 1) * LANES;
	uint32_t* sPtr = mem.colS + (size_t)(cx.slabOff + cs + 1) * LANES;
	int32_t nodeMax = L.scoreEnd;
	const uint64_t w = wStart + 1;
	const uint32_t* seqPtr = g.seq2 + (w >> 4) + 1;
	uint32_t seqWord = g.seq2[w >> 4];
	uint32_t shift = (uint32_t)(w & 15) * 2;
	for (uint32_t k = 1; k < len; k++)
	{
		// one loop for all lanes, refill test inside: per-word or per-32-base inner loops end at different columns per
		// lane (nodes start anywhere inside a sequence word) and were measured slower (divergence, exposed refill load)
		const uint32_t base = (seqWord >> shift) & 3u;
		shift += 2;
		if (shift == 32) { seqWord = *seqPtr++; shift = 0; }
		// match mask of this column's base from the slice's table (shared memory: the address depends only on the base)
		const uint64_t Eq = cx.eqTab[(size_t)base * LANES];
		const bool previousEq = ((prevMask >> base) & 1u) != 0;
		GaCol c;
		bool sbE = false;
		uint32_t eq0;
		uint32_t flags = GA_CF_PLAIN;   // the column is a plain word step from its left neighbour
		if (INPREV)
		{
			const uint32_t oldTiny = oldTinyNext;
			// software prefetch of the next column's previous-slice state (address known, value independent of this step)
			if (k + 1 < len) oldTinyNext = GA_TP(pcs + k + 1);
			const int32_t oldScore = ga_tiny_score(oldTiny);
			// row -1 score = min(left + 1, previous slice's end score); the flag says the latter attains it
			sbE = oldScore <= L.sbs + 1;
			bool needMerge;
			c = ga_next_col(Eq, L, LsbE, sbE, LsbE, previousEq, ga_tiny_row62(oldTinyLeft), oldScore, eq0, needMerge);
			if (needMerge) { ga_vertical_merge(c, oldScore); flags = 0; }
#ifdef GA_HOST_DEBUG
			{
				// the shortcut must equal the reference's step-then-merge bit for bit
				uint32_t e2; bool m2;
				GaCol ref = ga_next_col(Eq, L, LsbE, sbE, LsbE, previousEq, ga_tiny_row62(oldTinyLeft), 0x7fffffff, e2, m2);
				if (ref.sbs > oldScore) ga_vertical_merge(ref, oldScore);
				if (ref.VP != c.VP || ref.VN != c.VN || ref.sbs != c.sbs || ref.scoreEnd != c.scoreEnd) { fprintf(stderr, "vertical-merge shortcut mismatch\n"); abort(); }
			}
#endif
			oldTinyLeft = oldTiny;
		}
		else
		{
			bool needMerge;
			c = ga_next_col(Eq, L, LsbE, false, LsbE, previousEq, 0, 0x7fffffff, eq0, needMerge);
		}
		uint4 ra;
		ra.x = (uint32_t)c.VP; ra.y = (uint32_t)(c.VP >> 32); ra.z = (uint32_t)c.VN; ra.w = (uint32_t)(c.VN >> 32);
		*vvPtr = ra;
		*sPtr = (uint32_t)c.sbs | flags | (eq0 ? GA_CF_EQ0 : 0u);
		vvPtr += LANES;
		sPtr += LANES;
		GA_TC_ST(cs + k, ga_tiny_pack(c, sbE));
		if (c.scoreEnd < nodeMin) nodeMin = c.scoreEnd;
		if (SMALL && c.scoreEnd > nodeMax) nodeMax = c.scoreEnd;
		L = c;
		LsbE = sbE;
	}
	// the 16-bit tiny encoding holds scores up to GA_TINY_SPAN above the reference: beyond that the stream leaves small-band mode
	if (SMALL && nodeMax - cx.tinyRef > GA_TINY_SPAN) spanOverflow = true;
	return nodeMin;
}

#define GA_MAX_CACHED_IN 6

// Evaluate one band node whose band predecessors are final (the acyclic part of a band, in topological order): first column
// from its in-neighbours (or as a source), the rest by the word step.  Members of cyclic components go through
// ga_ex_calc_node instead.
template <int LANES, bool SMALL>
GA_DEV void ga_calc_node(const ga_graph_view& g, const ga_caps& caps, const GaLaneMem& mem, GaStreamState& st, const GaSliceCtx& cx, uint32_t slot)
{
	const uint32_t node = GA_HN(cx.nodeOff + slot, 0);
	const uint32_t cs = GA_HN(cx.nodeOff + slot, 1);
	const uint32_t len = GA_HN(cx.nodeOff + slot, 3);
	const uint64_t wStart = (uint64_t)mem.nWlo[(size_t)slot * LANES] | ((uint64_t)mem.nWhi[(size_t)slot * LANES] << 32);
	const uint32_t pcs = mem.nPcs[(size_t)slot * LANES];
	const bool inPrev = pcs != GA_NOT_IN_PREV;

	// ---- column 0 -------------------------------------------------------------------------------------------
	uint32_t seqWord = g.seq2[wStart >> 4];
	uint32_t base = (seqWord >> ((uint32_t)(wStart & 15) * 2)) & 3u;
	uint64_t Eq = base == 0 ? cx.BA : base == 1 ? cx.BC : base == 2 ? cx.BG : cx.BT;
	bool previousEq = cx.firstSlice ? inPrev : (base == cx.prevCharCode);
	const uint32_t oldTiny0 = inPrev ? GA_TP(pcs) : 0;
	uint32_t oldTinyNext = (inPrev && len > 1) ? GA_TP(pcs + 1) : 0;

	// in-neighbours that are in the current or the previous band: column index of their last column in the
	// current slab / in the previous tiny array (0xffffffff = absent)
	uint32_t inCur[GA_MAX_CACHED_IN], inPrevCol[GA_MAX_CACHED_IN], inSlot[GA_MAX_CACHED_IN];
	uint32_t nIn = 0;
	// row -1 score of the first column and its "exists" flag (forceComponentZeroRow, GraphAligner.h:1916-1989)
	int32_t sbs0 = inPrev ? ga_tiny_score(oldTiny0) : 0x7fffffff;
	const uint32_t eBegin = g.inOff[node], eEnd = g.inOff[node + 1];
	for (uint32_t e = eBegin; e < eEnd; e++)
	{
		uint32_t u = g.inAdj[e];
		int cu = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, u);
		int pu = ga_hash_find<LANES>(cx.hashPrev, cx.maskPrev, cx.stampPrev, u);
		if (cu < 0 && pu < 0) continue;
		uint32_t curCol = 0xffffffffu, prevCol = 0xffffffffu;
		if (cu >= 0)
		{
			const uint32_t ulen = GA_HN(cx.nodeOff + cu, 3);
			curCol = GA_HN(cx.nodeOff + cu, 1) + ulen - 1;
			{
				// small-band mode keeps the last column of every evaluated band node in shared memory
				int32_t v = (SMALL ? (int32_t)mem.lastS[(size_t)cu * LANES] : ga_col_load_sbs<LANES>(mem, cx.slabOff + curCol)) + 1;
				if (v < sbs0) sbs0 = v;
			}
		}
		if (pu >= 0)
		{
			prevCol = GA_HN(cx.pNodeOff + pu, 1) + GA_HN(cx.pNodeOff + pu, 3) - 1;
			{
				int32_t v = ga_tiny_score(GA_TP(prevCol)) + 1;
				if (v < sbs0) sbs0 = v;
			}
		}
		if (nIn < GA_MAX_CACHED_IN) { inCur[nIn] = curCol; inPrevCol[nIn] = prevCol; inSlot[nIn] = (uint32_t)cu; }
		nIn++;
	}
	GA_TLAP(st, 11);
	const bool sbE0 = inPrev && ga_tiny_score(oldTiny0) == sbs0;
	GaCol c0;
	c0.VP = 0; c0.VN = 0; c0.sbs = 0; c0.scoreEnd = 0;
	// A node whose only band in-neighbour is in this slice starts with a plain word step from that neighbour's last column:
	// it gets traceback masks and a link like any inner column (the walk then crosses the node border on the fast path)
	const bool single = nIn == 1 && inCur[0] != 0xffffffffu;
	uint32_t flags0 = 0;
	if (nIn > 0)
	{
		uint32_t k = 0;
		for (uint32_t e = eBegin; e < eEnd; e++)
		{
			uint32_t curCol, prevCol, curSlot;
			if (nIn <= GA_MAX_CACHED_IN)
			{
				if (k >= nIn) break;
				curCol = inCur[k];
				prevCol = inPrevCol[k];
				curSlot = inSlot[k];
			}
			else
			{
				// high in-degree: look the neighbour up again instead of caching
				uint32_t u = g.inAdj[e];
				int cu = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, u);
				int pu = ga_hash_find<LANES>(cx.hashPrev, cx.maskPrev, cx.stampPrev, u);
				if (cu < 0 && pu < 0) continue;
				curCol = cu >= 0 ? GA_HN(cx.nodeOff + cu, 1) + GA_HN(cx.nodeOff + cu, 3) - 1 : 0xffffffffu;
				curSlot = (uint32_t)cu;
				prevCol = pu >= 0 ? GA_HN(cx.pNodeOff + pu, 1) + GA_HN(cx.pNodeOff + pu, 3) - 1 : 0xffffffffu;
			}
			bool foundOneUp = prevCol != 0xffffffffu;
			uint32_t upTiny = foundOneUp ? GA_TP(prevCol) : 0;
			GaCol L;
			bool LsbE;
			uint64_t EqHere = Eq;
			if (curCol != 0xffffffffu)
			{
				if (SMALL)
				{
					L.VP = mem.lastVV[(size_t)(curSlot * 2) * LANES];
					L.VN = mem.lastVV[(size_t)(curSlot * 2 + 1) * LANES];
					L.sbs = (int32_t)mem.lastS[(size_t)curSlot * LANES];
				}
				else L = ga_col_load<LANES>(mem, cx.slabOff + curCol);
				uint32_t t = GA_TC(curCol);
				L.scoreEnd = ga_tiny_score(t);
				LsbE = (t & 4u) != 0;
			}
			else
			{
				// neighbour only in the previous band: synthetic source column from its end score (GraphAligner.h:1294-1301)
				int32_t es = ga_tiny_score(upTiny);
				L.VP = ~(uint64_t)0;
				L.VN = 0;
				L.sbs = es;
				L.scoreEnd = es + 64;
				LsbE = true;
				EqHere &= 1;
			}
			uint32_t eq0;
			bool needMerge;
			// single: the vertical merge rides on the step as in ga_node_columns (same recurrence, lower entry score)
			GaCol cand = ga_next_col(EqHere, L, LsbE, sbE0 && foundOneUp, foundOneUp, previousEq, ga_tiny_row62(upTiny), (single && inPrev) ? ga_tiny_score(oldTiny0) : 0x7fffffff, eq0, needMerge);
			if (single)
			{
				// the neighbour's last column sits right before this node's first one when the band lists the two back to back:
				// then the traceback walks across the node border as if the two nodes were one (GA_CF_LINK)
				if (!needMerge && curCol + 1 == cs) flags0 = GA_CF_LINK | (eq0 ? GA_CF_EQ0 : 0u);
				if (needMerge) ga_vertical_merge(cand, ga_tiny_score(oldTiny0));
#ifdef GA_HOST_DEBUG
				{
					uint32_t e2; bool m2;
					GaCol ref = ga_next_col(EqHere, L, LsbE, sbE0 && foundOneUp, foundOneUp, previousEq, ga_tiny_row62(upTiny), 0x7fffffff, e2, m2);
					if (inPrev && ref.sbs > ga_tiny_score(oldTiny0)) ga_vertical_merge(ref, ga_tiny_score(oldTiny0));
					if (ref.VP != cand.VP || ref.VN != cand.VN || ref.sbs != cand.sbs || ref.scoreEnd != cand.scoreEnd) { fprintf(stderr, "node-start shortcut mismatch\n"); abort(); }
				}
#endif
			}
			if (k == 0) c0 = cand;
			else c0 = ga_merge_cols(c0, cand);
			k++;
		}
		if (!single && inPrev && c0.sbs > ga_tiny_score(oldTiny0)) ga_vertical_merge(c0, ga_tiny_score(oldTiny0));
	}
	else
	{
		// source node (GraphAligner.h:1317-1347,1475-1488); a band node always has a band predecessor or is kept
		if (!inPrev) { st.status = GA_ERR_INTERNAL; return; }
		int32_t ps = ga_tiny_score(oldTiny0);
		uint64_t mismatch = 1;
		if (cx.firstSlice)
		{
			uint32_t m = st.aux[0] >> 4;
			mismatch = ((m >> base) & 1u) ? 0 : 1;
		}
		c0.VP = (~(uint64_t)1) | mismatch;
		c0.VN = 0;
		c0.scoreEnd = ps + 63 + (int32_t)mismatch;
		c0.sbs = ps;
	}
#ifdef GA_HOST_DEBUG
	if (c0.sbs != sbs0) { st.status = GA_ERR_INTERNAL; return; }
#endif
	GA_TLAP(st, 12);
	ga_col_store<LANES>(mem, cx.slabOff + cs, c0, flags0);
	GA_TC_ST(cs, ga_tiny_pack(c0, sbE0));

	// ---- columns 1 .. len-1 (GraphAligner.h:1532-1570) ------------------------------------------------------
	GA_TLAP(st, 2);
	int32_t nodeMin = c0.scoreEnd;
	if (len > 1)
	{
		// 4-bit set of graph bases that equal the read character just above the slice (exact compare,
		// GraphAligner.h:1540); on the first slice the flag is "node is in the previous band" instead
		const uint32_t prevMask = cx.firstSlice ? (inPrev ? 15u : 0u) : ((1u << cx.prevCharCode) & 15u);
		bool spanOverflow = false;
		if (inPrev) nodeMin = ga_node_columns<LANES, SMALL, true>(g, mem, cx, wStart, len, cs, pcs, prevMask, c0, sbE0, oldTiny0, oldTinyNext, nodeMin, spanOverflow);
		else nodeMin = ga_node_columns<LANES, SMALL, false>(g, mem, cx, wStart, len, cs, pcs, prevMask, c0, sbE0, oldTiny0, oldTinyNext, nodeMin, spanOverflow);
		if (spanOverflow) { st.status = GA_ERR_COL_OVERFLOW; return; }
	}
	else if (SMALL && c0.scoreEnd - cx.tinyRef > GA_TINY_SPAN) { st.status = GA_ERR_COL_OVERFLOW; return; }
	GA_HN(cx.nodeOff + slot, 2) = (uint32_t)nodeMin;
	if (SMALL)
	{
		mem.lastVV[(size_t)(slot * 2) * LANES] = c0.VP;
		mem.lastVV[(size_t)(slot * 2 + 1) * LANES] = c0.VN;
		mem.lastS[(size_t)slot * LANES] = (uint32_t)c0.sbs;
	}
}


// ================================================================================================================
// Exact emulation of the reference's evaluation of a CYCLIC band component (GraphAligner.h:1349-1427, 1457-1573,
// 2360-2420; WordSlice.h:361-510).  Cell values of a cyclic component are a fix point and could be had any way, but
// two things the reference feeds into later slices are artefacts of HOW it iterates: the per-node minimum it stores
// is the one of the LAST calculateNode call on that node (which covers only the columns that call confirmed), and the
// order of tied minimum cells follows its work list.  So for these components the confirmedRows schedule is replayed.
// ================================================================================================================
struct GaExCol
{
	uint64_t VP, VN;
	int32_t sbs, scoreEnd;
	int32_t rows;      // confirmedRows.rows
	bool partial;      // confirmedRows.partial
	bool sbE;          // scoreBeforeExists
};

GA_DEV bool ga_conf_gt(int ra, bool pa, int rb, bool pb) { return ra > rb || (ra == rb && pa && !pb); }   // RowConfirmation::operator>
GA_DEV bool ga_conf_eq(int ra, bool pa, int rb, bool pb) { return ra == rb && pa == pb; }

// getNextSlice with the confirmedRows bookkeeping (GraphAligner.h:1349-1427).  Shift counts follow x86 (masked to 6 bits),
// which is what the reference's build does with rows == 64 and rows == 0.
GA_DEV GaExCol ga_ex_next(uint64_t Eq, GaExCol slice, bool upInside, bool upleftInside, bool diagInside, bool previousEq, int32_t prevScoreEnd, uint32_t prevVP63, uint32_t prevVN63)
{
	const int32_t oldValue = slice.sbs;
	const uint64_t confirmedMask = (uint64_t)1 << (slice.rows & 63);
	const uint64_t prevConfirmedMask = (uint64_t)1 << ((slice.rows - 1) & 63);
	bool confirmOneMore = false;
	if (!slice.sbE) Eq &= ~(uint64_t)1;
	slice.sbE = upInside;
	if (!diagInside) Eq &= ~(uint64_t)1;
	if (!upleftInside) slice.sbs += 1;
	else
	{
		int32_t d = prevScoreEnd - (int32_t)prevVP63 + (int32_t)prevVN63 + (previousEq ? 0 : 1);
		slice.sbs = slice.sbs + 1 < d ? slice.sbs + 1 : d;
	}
	const int32_t hin = slice.sbs - oldValue;
	uint64_t Xv = Eq | slice.VN;
	if (hin < 0) Eq |= 1;
	uint64_t Xh = (((Eq & slice.VP) + slice.VP) ^ slice.VP) | Eq;
	uint64_t Ph = slice.VN | ~(Xh | slice.VP);
	uint64_t Mh = slice.VP & Xh;
	int32_t diagonalDiff = hin;
	if (slice.rows > 0) diagonalDiff = ((Ph & prevConfirmedMask) ? 1 : 0) - ((Mh & prevConfirmedMask) ? 1 : 0);
	if (slice.rows > 0 && (Mh & prevConfirmedMask)) confirmOneMore = true;
	else if (slice.rows == 0 && hin == -1) confirmOneMore = true;
	const uint64_t lastBitMask = (uint64_t)1 << 63;
	if (Ph & lastBitMask) slice.scoreEnd += 1;
	else if (Mh & lastBitMask) slice.scoreEnd -= 1;
	if (slice.partial && (~Ph & confirmedMask)) confirmOneMore = true;
	Ph <<= 1;
	Mh <<= 1;
	if (hin < 0) Mh |= 1; else if (hin > 0) Ph |= 1;
	slice.VP = Mh | ~(Xv | Ph);
	slice.VN = Ph & Xv;
	diagonalDiff += ((slice.VP & confirmedMask) ? 1 : 0) - ((slice.VN & confirmedMask) ? 1 : 0);
	if (diagonalDiff <= 0) confirmOneMore = true;
	else if (slice.VN & confirmedMask) confirmOneMore = true;
	if (confirmOneMore)
	{
		if (slice.rows + 1 <= 64) slice.rows += 1;
		slice.partial = false;
	}
	else if (!slice.partial && slice.rows < 64) slice.partial = true;
	return slice;
}

// position of the rank-th set bit of a 128-bit value (low word first), WordSlice.h:46-98
GA_DEV int ga_bit_position64(uint64_t number, int rank)
{
	int total = (int)GA_POPC(number);
	if (rank >= total) return 64 + (rank - total);
	for (int i = 0; i < rank; i++) number &= number - 1;
	return (int)GA_CTZ(number);
}
GA_DEV int ga_bit_position(uint64_t low, uint64_t high, int rank)
{
	int result = ga_bit_position64(low, rank);
	if (result < 64) return result;
	return 64 + ga_bit_position64(high, result - 64);
}
// bits of x on the even positions, bits of y on the odd ones (WordSlice.h:111-131)
GA_DEV uint64_t ga_interleave(uint64_t x, uint64_t y)
{
	x = (x | (x << 16)) & 0x0000FFFF0000FFFFull; x = (x | (x << 8)) & 0x00FF00FF00FF00FFull; x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
	x = (x | (x << 2)) & 0x3333333333333333ull; x = (x | (x << 1)) & 0x5555555555555555ull;
	y = (y | (y << 16)) & 0x0000FFFF0000FFFFull; y = (y | (y << 8)) & 0x00FF00FF00FF00FFull; y = (y | (y << 4)) & 0x0F0F0F0F0F0F0F0Full;
	y = (y | (y << 2)) & 0x3333333333333333ull; y = (y | (y << 1)) & 0x5555555555555555ull;
	return x | (y << 1);
}

// confirmedRowsInMerged, WordSlice.h:423-510
GA_DEV void ga_ex_conf_merged(GaExCol left, GaExCol right, int32_t& rowsOut, bool& partialOut)
{
	if (ga_conf_eq(left.rows, left.partial, right.rows, right.partial)) { rowsOut = left.rows; partialOut = left.partial; return; }
	if (ga_conf_gt(right.rows, right.partial, left.rows, left.partial)) { GaExCol t = left; left = right; right = t; }
	int32_t leftScore = left.sbs, rightScore = right.sbs;
	const uint64_t confirmedMask = ~(~(uint64_t)0 << (right.rows & 63));   // rows == 64 shifts by 0 on x86, like the reference's build
	leftScore += (int32_t)GA_POPC(left.VP & confirmedMask) - (int32_t)GA_POPC(left.VN & confirmedMask);
	rightScore += (int32_t)GA_POPC(right.VP & confirmedMask) - (int32_t)GA_POPC(right.VN & confirmedMask);
	if (right.rows == left.rows)
	{
		const uint64_t mask = (uint64_t)1 << (left.rows & 63);
		rightScore -= 1;
		if (!(left.VP & mask)) leftScore -= 1;
		rowsOut = left.rows;
		partialOut = leftScore <= rightScore;
		return;
	}
	const uint64_t premask = (uint64_t)1 << (right.rows & 63);
	leftScore += (left.VP & premask) ? 1 : 0;
	leftScore -= (left.VN & premask) ? 1 : 0;
	if (!(right.partial && (right.VP & premask))) rightScore -= 1;
	if (leftScore == rightScore + 1) { rowsOut = right.rows; partialOut = true; return; }
	if (leftScore > rightScore + 1) { rowsOut = right.rows; partialOut = right.partial; return; }
	if (left.rows > right.rows + 1)
	{
		uint64_t partiallyConfirmedMask = 0;
		if (left.rows < 64) partiallyConfirmedMask = ~(uint64_t)0 << left.rows;
		partiallyConfirmedMask = ~partiallyConfirmedMask;
		partiallyConfirmedMask &= ~(uint64_t)0 << (right.rows + 1);
		const uint64_t low = left.VP & partiallyConfirmedMask;
		const uint64_t high = ~left.VN & partiallyConfirmedMask;
		const uint64_t mortonLow = ga_interleave(low & 0xFFFFFFFFull, high & 0xFFFFFFFFull);
		const uint64_t mortonHigh = ga_interleave(low >> 32, high >> 32);
		const int pos = ga_bit_position(mortonLow, mortonHigh, rightScore - leftScore);
		if (pos / 2 < left.rows)
		{
			const int nextpos = ga_bit_position(mortonLow, mortonHigh, rightScore - leftScore + 1);
			rowsOut = pos / 2;
			partialOut = nextpos / 2 > pos / 2;
			return;
		}
		leftScore += (int32_t)GA_POPC(left.VP & partiallyConfirmedMask) - (int32_t)GA_POPC(left.VN & partiallyConfirmedMask);
		rightScore -= left.rows - right.rows - 1;
	}
	if (!left.partial) { rowsOut = left.rows; partialOut = left.partial; return; }
	const uint64_t postmask = (uint64_t)1 << (left.rows & 63);
	rightScore -= 1;
	if (left.VP & postmask)
	{
		if (leftScore <= rightScore) { rowsOut = left.rows; partialOut = left.partial; return; }
	}
	else
	{
		rowsOut = left.rows;
		partialOut = left.partial;
		return;
	}
	rowsOut = left.rows;
	partialOut = false;
}

// WordSlice::mergeWith -> mergeTwoSlices, WordSlice.h:202-206,361-421 (values by exact minimum, flags by the reference's rules)
GA_DEV GaExCol ga_ex_merge(GaExCol left, GaExCol right)
{
	if (left.sbs > right.sbs) { GaExCol t = left; left = right; right = t; }
	GaExCol result;
	ga_ex_conf_merged(left, right, result.rows, result.partial);
	GaCol a, b;
	a.VP = left.VP; a.VN = left.VN; a.sbs = left.sbs; a.scoreEnd = left.scoreEnd;
	b.VP = right.VP; b.VN = right.VN; b.sbs = right.sbs; b.scoreEnd = right.scoreEnd;
	GaCol m = ga_merge_cols(a, b);
	result.VP = m.VP; result.VN = m.VN; result.sbs = m.sbs; result.scoreEnd = m.scoreEnd;
	if (left.sbs < right.sbs) result.sbE = left.sbE;
	else result.sbE = left.sbE || right.sbE;   // equal after the swap
	return result;
}

template <int LANES, bool SMALL>
GA_DEV GaExCol ga_ex_load(const GaLaneMem& mem, const GaSliceCtx& cx, uint32_t col)
{
	GaCol c = ga_col_load<LANES>(mem, cx.slabOff + col);
	uint32_t t = GA_TC(col);
	uint32_t cf = mem.conf[(size_t)col * LANES];
	GaExCol e;
	e.VP = c.VP; e.VN = c.VN; e.sbs = c.sbs; e.scoreEnd = c.scoreEnd;
	e.rows = (int32_t)(cf & 0xffu);
	e.partial = (cf & 0x100u) != 0;
	e.sbE = (t & 4u) != 0;
	return e;
}

template <int LANES, bool SMALL>
GA_DEV void ga_ex_store(const GaLaneMem& mem, const GaSliceCtx& cx, uint32_t col, const GaExCol& e)
{
	GaCol c;
	c.VP = e.VP; c.VN = e.VN; c.sbs = e.sbs; c.scoreEnd = e.scoreEnd;
	ga_col_store<LANES>(mem, cx.slabOff + col, c, 0);
	GA_TC_ST(col, ga_tiny_pack(c, e.sbE));
	mem.conf[(size_t)col * LANES] = (uint32_t)e.rows | (e.partial ? 0x100u : 0u);
}

// calculateNode for a member of a cyclic component (GraphAligner.h:1457-1573).  Returns the minimum scoreEnd over the
// columns this call fully confirmed (INT_MAX if none) and the last such column attaining it.
template <int LANES, bool SMALL>
GA_DEV int32_t ga_ex_calc_node(const ga_graph_view& g, const ga_caps& caps, const GaLaneMem& mem, GaStreamState& st, const GaSliceCtx& cx, uint32_t slot, uint32_t& lastMinCol)
{
	const uint32_t node = GA_HN(cx.nodeOff + slot, 0);
	const uint32_t cs = GA_HN(cx.nodeOff + slot, 1);
	const uint32_t len = GA_HN(cx.nodeOff + slot, 3);
	const uint64_t wStart = (uint64_t)mem.nWlo[(size_t)slot * LANES] | ((uint64_t)mem.nWhi[(size_t)slot * LANES] << 32);
	const uint32_t pcs = mem.nPcs[(size_t)slot * LANES];
	const bool inPrev = pcs != GA_NOT_IN_PREV;
	int32_t minScore = 0x7fffffff;
	lastMinCol = 0xffffffffu;
	GaExCol cur0 = ga_ex_load<LANES, SMALL>(mem, cx, cs);
	if (cur0.rows == 64) return minScore;
	const int32_t oldRows0 = cur0.rows;
	const bool oldPartial0 = cur0.partial;
	uint32_t base = ga_base(g, wStart);
	uint64_t Eq = base == 0 ? cx.BA : base == 1 ? cx.BC : base == 2 ? cx.BG : cx.BT;
	bool previousEq = cx.firstSlice ? inPrev : (base == cx.prevCharCode);
	// getNodeStartSlice, GraphAligner.h:1270-1315
	GaExCol res;
	bool foundOne = false;
	for (uint32_t e = g.inOff[node], eEnd = g.inOff[node + 1]; e < eEnd; e++)
	{
		uint32_t u = g.inAdj[e];
		int cu = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, u);
		int pu = ga_hash_find<LANES>(cx.hashPrev, cx.maskPrev, cx.stampPrev, u);
		if (cu < 0 && pu < 0) continue;
		uint64_t EqHere = Eq;
		const bool foundOneUp = pu >= 0;
		uint32_t upTiny = 0;
		if (foundOneUp) upTiny = GA_TP(GA_HN(cx.pNodeOff + pu, 1) + GA_HN(cx.pNodeOff + pu, 3) - 1);
		GaExCol previous;
		if (cu >= 0) previous = ga_ex_load<LANES, SMALL>(mem, cx, GA_HN(cx.nodeOff + cu, 1) + GA_HN(cx.nodeOff + cu, 3) - 1);
		else
		{
			const int32_t es = ga_tiny_score(upTiny);
			previous.VP = ~(uint64_t)0; previous.VN = 0; previous.scoreEnd = es + 64; previous.sbs = es; previous.rows = 64; previous.partial = false; previous.sbE = true;
			EqHere &= 1;
		}
		GaExCol here = ga_ex_next(EqHere, previous, cur0.sbE, cur0.sbE && foundOneUp, foundOneUp, previousEq, ga_tiny_score(upTiny), upTiny & 1u, (upTiny >> 1) & 1u);
		if (!foundOne) { res = here; foundOne = true; }
		else res = ga_ex_merge(res, here);
	}
	uint32_t oldTiny = inPrev ? GA_TP(pcs) : 0;
	if (!foundOne)
	{
		// source node downstream of nothing (GraphAligner.h:1317-1347,1475-1488): final at once
		if (!inPrev) { st.status = GA_ERR_INTERNAL; return minScore; }
		const int32_t ps = ga_tiny_score(oldTiny);
		uint64_t mismatch = 1;
		if (cx.firstSlice) mismatch = (((st.aux[0] >> 4) >> base) & 1u) ? 0 : 1;
		res.VP = (~(uint64_t)1) | mismatch; res.VN = 0; res.scoreEnd = ps + 63 + (int32_t)mismatch; res.sbs = ps;
		res.rows = 64; res.partial = false; res.sbE = true;
	}
	else if (inPrev && res.sbs > ga_tiny_score(oldTiny))
	{
		GaExCol mergable;
		mergable.VP = ~(uint64_t)0; mergable.VN = 0; mergable.sbs = ga_tiny_score(oldTiny); mergable.scoreEnd = mergable.sbs + 64; mergable.rows = 64; mergable.partial = false; mergable.sbE = true;
		res = ga_ex_merge(res, mergable);
	}
	ga_ex_store<LANES, SMALL>(mem, cx, cs, res);
	if (res.rows == 64 && res.scoreEnd < minScore) minScore = res.scoreEnd;
	if (res.rows == 64 && res.scoreEnd == minScore) lastMinCol = 0;
	if (ga_conf_eq(res.rows, res.partial, oldRows0, oldPartial0)) return minScore;
	GaExCol leftCol = res;
	uint32_t oldTinyLeft = oldTiny;
	for (uint32_t k = 1; k < len; k++)
	{
		GaExCol c = ga_ex_load<LANES, SMALL>(mem, cx, cs + k);
		if (c.rows == 64) return minScore;
		const int32_t oldRows = c.rows;
		const bool oldPartial = c.partial;
		base = ga_base(g, wStart + k);
		Eq = base == 0 ? cx.BA : base == 1 ? cx.BC : base == 2 ? cx.BG : cx.BT;
		previousEq = cx.firstSlice ? inPrev : (base == cx.prevCharCode);
		oldTiny = inPrev ? GA_TP(pcs + k) : 0;
		GaExCol n = ga_ex_next(Eq, leftCol, c.sbE, c.sbE, leftCol.sbE, previousEq, ga_tiny_score(oldTinyLeft), oldTinyLeft & 1u, (oldTinyLeft >> 1) & 1u);
		if (inPrev && n.sbs > ga_tiny_score(oldTiny))
		{
			GaExCol mergable;
			mergable.VP = ~(uint64_t)0; mergable.VN = 0; mergable.sbs = ga_tiny_score(oldTiny); mergable.scoreEnd = mergable.sbs + 64; mergable.rows = 64; mergable.partial = false; mergable.sbE = true;
			n = ga_ex_merge(n, mergable);
		}
		ga_ex_store<LANES, SMALL>(mem, cx, cs + k, n);
		if (n.rows == 64 && n.scoreEnd < minScore) minScore = n.scoreEnd;
		if (n.rows == 64 && n.scoreEnd == minScore) lastMinCol = k;
		if (ga_conf_eq(n.rows, n.partial, oldRows, oldPartial)) return minScore;
		leftCol = n;
		oldTinyLeft = oldTiny;
	}
	return minScore;
}

// Row -1 scores for a cyclic block of band nodes (the slots listed in order[from..to)) by shortest paths over
// the block (forceComponentZeroRow, GraphAligner.h:1903-1995), then reset every column to the all-ones ramp.
template <int LANES, bool SMALL>
GA_DEV void ga_force_block(const ga_graph_view& g, const ga_caps& caps, const GaLaneMem& mem, GaStreamState& st, const GaSliceCtx& cx, uint32_t from, uint32_t to, uint32_t comp)
{
	const int32_t INF = 0x1fffffff;   // fits the score field of a column record (GA_CF_SCORE_MASK)
	uint32_t heapN = 0;
	// the component's members are emit[from..to); cmpOf[slot] == comp tests membership
	for (uint32_t q = from; q < to; q++)
	{
		uint32_t slot = mem.emit[(size_t)q * LANES];
		uint32_t node = GA_HN(cx.nodeOff + slot, 0);
		uint32_t cs = GA_HN(cx.nodeOff + slot, 1);
		uint32_t len = GA_HN(cx.nodeOff + slot, 3);
		uint32_t pcs = mem.nPcs[(size_t)slot * LANES];
		bool inPrev = pcs != GA_NOT_IN_PREV;
		int32_t s0 = inPrev ? ga_tiny_score(GA_TP(pcs)) : INF;
		for (uint32_t e = g.inOff[node]; e < g.inOff[node + 1]; e++)
		{
			uint32_t u = g.inAdj[e];
			int cu = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, u);
			int pu = ga_hash_find<LANES>(cx.hashPrev, cx.maskPrev, cx.stampPrev, u);
			if (cu >= 0 && (mem.cmpOf[(size_t)cu * LANES] & 0x3fffffffu) != comp)
			{
				uint32_t ucol = GA_HN(cx.nodeOff + cu, 1) + GA_HN(cx.nodeOff + cu, 3) - 1;
				int32_t v = ga_col_load_sbs<LANES>(mem, cx.slabOff + ucol) + 1;
				if (v < s0) s0 = v;
			}
			if (pu >= 0)
			{
				uint32_t ucol = GA_HN(cx.pNodeOff + pu, 1) + GA_HN(cx.pNodeOff + pu, 3) - 1;
				int32_t v = ga_tiny_score(GA_TP(ucol)) + 1;
				if (v < s0) s0 = v;
			}
		}
		int32_t v = s0;
		for (uint32_t k = 0; k < len; k++)
		{
			if (k > 0 && v < INF)
			{
				v = v + 1;
				if (inPrev)
				{
					int32_t o = ga_tiny_score(GA_TP(pcs + k));
					if (o < v) v = o;
				}
			}
			ga_col_store_sbs<LANES>(mem, cx.slabOff + cs + k, v);
		}
		if (v < INF)
		{
			for (uint32_t e = g.outOff[node]; e < g.outOff[node + 1]; e++)
			{
				int cu = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, g.outAdj[e]);
				if (cu < 0 || (mem.cmpOf[(size_t)cu * LANES] & 0x3fffffffu) != comp) continue;
				if (heapN >= caps.maxQueue) { st.status = GA_ERR_QUEUE_OVERFLOW; return; }
				ga_heap_push<LANES>(mem.heap, heapN, ((uint64_t)(uint32_t)(v + 1) << 32) | (uint32_t)cu);
			}
		}
	}
	while (heapN > 0)
	{
		uint64_t top = ga_heap_pop<LANES>(mem.heap, heapN);
		int32_t score = (int32_t)(top >> 32);
		uint32_t slot = (uint32_t)top;
		uint32_t node = GA_HN(cx.nodeOff + slot, 0);
		uint32_t cs = GA_HN(cx.nodeOff + slot, 1);
		uint32_t len = GA_HN(cx.nodeOff + slot, 3);
		bool endUpdated = true;
		for (uint32_t k = 0; k < len; k++)
		{
			if (ga_col_load_sbs<LANES>(mem, cx.slabOff + cs + k) <= score) { endUpdated = false; break; }
			ga_col_store_sbs<LANES>(mem, cx.slabOff + cs + k, score);
			score++;
		}
		if (!endUpdated) continue;
		for (uint32_t e = g.outOff[node]; e < g.outOff[node + 1]; e++)
		{
			int cu = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, g.outAdj[e]);
			if (cu < 0 || (mem.cmpOf[(size_t)cu * LANES] & 0x3fffffffu) != comp) continue;
			if (heapN >= caps.maxQueue) { st.status = GA_ERR_QUEUE_OVERFLOW; return; }
			ga_heap_push<LANES>(mem.heap, heapN, ((uint64_t)(uint32_t)score << 32) | (uint32_t)cu);
		}
	}
	// reset columns to {VP = all ones, VN = 0, scoreEnd = sbs + 64, confirmedRows = 0, scoreBeforeExists} (GraphAligner.h:1981-1993)
	for (uint32_t q = from; q < to; q++)
	{
		uint32_t slot = mem.emit[(size_t)q * LANES];
		uint32_t cs = GA_HN(cx.nodeOff + slot, 1);
		uint32_t len = GA_HN(cx.nodeOff + slot, 3);
		uint32_t pcs = mem.nPcs[(size_t)slot * LANES];
		bool inPrev = pcs != GA_NOT_IN_PREV;
		for (uint32_t k = 0; k < len; k++)
		{
			GaCol c;
			c.VP = ~(uint64_t)0;
			c.VN = 0;
			c.sbs = ga_col_load_sbs<LANES>(mem, cx.slabOff + cs + k);
			if (c.sbs >= INF) { st.status = GA_ERR_INTERNAL; return; }
			c.scoreEnd = c.sbs + 64;
			bool sbE = inPrev && ga_tiny_score(GA_TP(pcs + k)) == c.sbs;
			ga_col_store<LANES>(mem, cx.slabOff + cs + k, c, 0);
			GA_TC_ST(cs + k, ga_tiny_pack(c, sbE));
			mem.conf[(size_t)(cs + k) * LANES] = 0;
		}
	}
}


// Tarjan over the whole band in the reference's visiting order (band order, outNeighbors order, GraphAligner.h:1759-1856).
// Fills emit[] with the band slots in emission order (the members of a component are contiguous, in the order the
// reference's component vector holds them) and cmpOf[slot] with the component number (emission order).  Returns the
// number of components.  Scratch: indeg = DFS index | on-stack bit, order = low link, uorder = Tarjan stack,
// unext = call-stack slots, ubkt = call-stack edge cursors.  (The hash table of the slice resolves neighbours.)
template <int LANES, bool SMALL>
GA_DEV uint32_t ga_tarjan_components(const ga_graph_view& g, const ga_caps& caps, const GaLaneMem& mem, const GaSliceCtx& cx)
{
	const uint32_t nNodes = cx.nNodes;
	const uint32_t ONSTACK = 0x80000000u;
	for (uint32_t i = 0; i < nNodes; i++) mem.indeg[(size_t)i * LANES] = 0;
	uint32_t counter = 0, tstack = 0, emitted = 0, nComp = 0;
	for (uint32_t root = 0; root < nNodes; root++)
	{
		if (mem.indeg[(size_t)root * LANES] != 0) continue;
		uint32_t depth = 0;
		mem.unext[0] = root;
		mem.ubkt[0] = g.outOff[GA_HN(cx.nodeOff + root, 0)];
		counter++;
		mem.indeg[(size_t)root * LANES] = counter | ONSTACK;
		mem.order[(size_t)root * LANES] = counter;
		mem.uorder[(size_t)(tstack++) * LANES] = root;
		while (true)
		{
			uint32_t slot = mem.unext[(size_t)depth * LANES];
			uint32_t node = GA_HN(cx.nodeOff + slot, 0);
			uint32_t e = mem.ubkt[(size_t)depth * LANES];
			if (e < g.outOff[node + 1])
			{
				mem.ubkt[(size_t)depth * LANES] = e + 1;
				int nb = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, g.outAdj[e]);
				if (nb < 0) continue;
				uint32_t mark = mem.indeg[(size_t)nb * LANES];
				if (mark == 0)
				{
					depth++;
					mem.unext[(size_t)depth * LANES] = (uint32_t)nb;
					mem.ubkt[(size_t)depth * LANES] = g.outOff[GA_HN(cx.nodeOff + nb, 0)];
					counter++;
					mem.indeg[(size_t)nb * LANES] = counter | ONSTACK;
					mem.order[(size_t)nb * LANES] = counter;
					mem.uorder[(size_t)(tstack++) * LANES] = (uint32_t)nb;
				}
				else if (mark & ONSTACK)
				{
					uint32_t idx = mark & ~ONSTACK;
					if (idx < mem.order[(size_t)slot * LANES]) mem.order[(size_t)slot * LANES] = idx;
				}
				continue;
			}
			uint32_t low = mem.order[(size_t)slot * LANES];
			if (low == (mem.indeg[(size_t)slot * LANES] & ~ONSTACK))
			{
				while (true)
				{
					uint32_t back = mem.uorder[(size_t)(--tstack) * LANES];
					mem.indeg[(size_t)back * LANES] &= ~ONSTACK;
					mem.emit[(size_t)(emitted++) * LANES] = back;
					mem.cmpOf[(size_t)back * LANES] = nComp;
					if (back == slot) break;
				}
				nComp++;
			}
			if (depth == 0) break;
			depth--;
			uint32_t parent = mem.unext[(size_t)depth * LANES];
			if (low < mem.order[(size_t)parent * LANES]) mem.order[(size_t)parent * LANES] = low;
		}
	}
	return nComp;
}

struct GaSliceResult
{
	int32_t minScore;
	double hmmC, hmmF;
	bool correctFromCorrect, falseFromCorrect, currentlyCorrect;
};

// One slice for one stream, after band selection: topological pass (Kahn) over the acyclic part, fix-point
// sweeps over what is left (cyclic components and everything downstream of them), slice minimum, HMM step.
// Returns false on error; the slice's minimum and HMM step come back in res (the caller applies the stop / ramp rules).
template <int LANES, bool SMALL>
GA_DEV bool ga_fill_slice(const ga_graph_view& g, const ga_caps& caps, const GaHmmTables& hmm, const GaLaneMem& mem, GaStreamState& st, GaSliceCtx& cx, uint32_t ncols, GaSliceResult& res)
{
	const uint32_t nc = cx.nNodes;
	// Peq words for the 64 read characters of this slice (GraphAligner.h:2338-2351), precomputed by ga_peq_kernel
	{
		uint4 a = mem.peq[(size_t)cx.s * 2], b = mem.peq[(size_t)cx.s * 2 + 1];
		cx.BA = (uint64_t)a.x | ((uint64_t)a.y << 32);
		cx.BC = (uint64_t)a.z | ((uint64_t)a.w << 32);
		cx.BG = (uint64_t)b.x | ((uint64_t)b.y << 32);
		cx.BT = (uint64_t)b.z | ((uint64_t)b.w << 32);
		cx.eqTab = mem.eqTab;
		cx.eqTab[0] = cx.BA; cx.eqTab[LANES] = cx.BC; cx.eqTab[2 * LANES] = cx.BG; cx.eqTab[3 * LANES] = cx.BT;
		cx.prevCharCode = cx.s > 0 ? (st.aux[cx.s] & 7u) : 4;
		cx.firstSlice = cx.s == 0;
	}
	GA_TLAP(st, 13);
	// in-degrees inside the band
	uint32_t ready = 0;   // order[0..ready) = nodes whose predecessors are all evaluated (FIFO)
	for (uint32_t slot = 0; slot < nc; slot++)
	{
		uint32_t node = GA_HN(cx.nodeOff + slot, 0);
		uint32_t d = 0;
		for (uint32_t e = g.inOff[node], eEnd = g.inOff[node + 1]; e < eEnd; e++)
		{
			if (ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, g.inAdj[e]) >= 0) d++;
		}
		mem.indeg[(size_t)slot * LANES] = d;
		if (d == 0) mem.order[(size_t)(ready++) * LANES] = slot;
	}
	uint32_t done = 0;
	GA_TLAP(st, 1);
	while (done < ready)
	{
		uint32_t slot = mem.order[(size_t)(done++) * LANES];
		ga_calc_node<LANES, SMALL>(g, caps, mem, st, cx, slot);
		GA_TLAP(st, 3);
		if (st.status != GA_OK) return false;
		uint32_t node = GA_HN(cx.nodeOff + slot, 0);
		for (uint32_t e = g.outOff[node], eEnd = g.outOff[node + 1]; e < eEnd; e++)
		{
			int cu = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, g.outAdj[e]);
			if (cu < 0) continue;
			uint32_t d = mem.indeg[(size_t)cu * LANES] - 1;
			mem.indeg[(size_t)cu * LANES] = d;
			if (d == 0) mem.order[(size_t)(ready++) * LANES] = (uint32_t)cu;
		}
		GA_TLAP(st, 4);
	}
	int32_t minScore = 0x7fffffff;
	uint32_t lastMinSlot = 0xffffffffu, lastMinCol = 0;
	if (done == nc)
	{
		// acyclic band: slice minimum over the per-node minima (GraphAligner.h:2375,2410-2418); the order of tied cells is
		// resolved later, only for the slice the traceback starts from (ga_first_emitted_min_node)
		for (uint32_t slot = 0; slot < nc; slot++)
		{
			int32_t nodeMin = (int32_t)GA_HN(cx.nodeOff + slot, 2);
			if (nodeMin < minScore) minScore = nodeMin;
		}
	}
	else
	{
		// small-band mode covers acyclic bands only: the stream is re-run with the general layout
		if (SMALL) { st.status = GA_ERR_NODE_OVERFLOW; return false; }
		// The band holds a cycle.  Components in the reference's order (reverse Tarjan emission = topological); acyclic
		// ones that Kahn's pass did not reach are evaluated like any other node, cyclic ones replay the reference's
		// confirmedRows work list.  The slice minimum and its last tied cell are tracked in evaluation order.
		st.cyclicSlices++;
		// mark the nodes Kahn's pass evaluated (indeg is about to be reused by Tarjan)
		for (uint32_t slot = 0; slot < nc; slot++) mem.wl[(size_t)slot * LANES] = 0;
		for (uint32_t q = 0; q < done; q++) mem.wl[(size_t)mem.order[(size_t)q * LANES] * LANES] = 1;
		// wl doubles as the "already evaluated" flag array until the work list needs it: copy the flags into cmpOf's top bit
		const uint32_t nComp = ga_tarjan_components<LANES, SMALL>(g, caps, mem, cx);
		// columns outside cyclic components are final when a component reads them: confirmedRows = 64
		for (uint32_t c = 0; c < ncols; c++) mem.conf[(size_t)c * LANES] = 64;
		for (uint32_t slot = 0; slot < nc; slot++)
		{
			if (mem.wl[(size_t)slot * LANES]) mem.cmpOf[(size_t)slot * LANES] |= 0x40000000u;
		}
		bool sawCyclic = false;
		uint32_t compEnd = nc;   // emit[compStart..compEnd) = current component, walking emission order backwards
		for (uint32_t ci = nComp; ci-- > 0;)
		{
			uint32_t compStart = compEnd;
			while (compStart > 0 && (mem.cmpOf[(size_t)mem.emit[(size_t)(compStart - 1) * LANES] * LANES] & 0x3fffffffu) == ci) compStart--;
			const uint32_t firstSlot = mem.emit[(size_t)compStart * LANES];
			bool cyclic = compEnd - compStart > 1;
			if (!cyclic)
			{
				const uint32_t node = GA_HN(cx.nodeOff + firstSlot, 0);
				for (uint32_t e = g.outOff[node], eEnd = g.outOff[node + 1]; e < eEnd; e++) cyclic = cyclic || g.outAdj[e] == node;
			}
			// once a cyclic component has been evaluated, everything after it may read columns the reference left
			// partly confirmed, so it replays the reference's path as well
			const bool exact = cyclic || sawCyclic;
			sawCyclic = sawCyclic || cyclic;
			if (!exact)
			{
				if (!(mem.cmpOf[(size_t)firstSlot * LANES] & 0x40000000u))
				{
					ga_calc_node<LANES, SMALL>(g, caps, mem, st, cx, firstSlot);
					if (st.status != GA_OK) return false;
				}
				const int32_t nodeMin = (int32_t)GA_HN(cx.nodeOff + firstSlot, 2);
				if (nodeMin <= minScore)
				{
					// minScoreIndex gets this node's tied columns in ascending order: the last one is the highest
					const uint32_t cs = GA_HN(cx.nodeOff + firstSlot, 1), len = GA_HN(cx.nodeOff + firstSlot, 3);
					for (uint32_t k = 0; k < len; k++)
					{
						if (ga_tiny_score(GA_TC(cs + k)) == nodeMin) lastMinCol = k;
					}
					minScore = nodeMin;
					lastMinSlot = firstSlot;
				}
			}
			else
			{
				// forceComponentZeroRow, then the UniqueQueue work list (GraphAligner.h:2360-2420, UniqueQueue.h)
				ga_force_block<LANES, SMALL>(g, caps, mem, st, cx, compStart, compEnd, ci);
				if (st.status != GA_OK) return false;
				uint32_t wlN = 0;
				const uint32_t INQ = 0x80000000u;
				for (uint32_t q = compStart; q < compEnd; q++)
				{
					const uint32_t slot = mem.emit[(size_t)q * LANES];
					mem.wl[(size_t)(wlN++) * LANES] = slot;
					mem.cmpOf[(size_t)slot * LANES] |= INQ;
				}
				uint32_t guard = 0;
				while (wlN > 0)
				{
					const uint32_t slot = mem.wl[(size_t)(--wlN) * LANES];
					mem.cmpOf[(size_t)slot * LANES] &= ~INQ;
					const uint32_t cs = GA_HN(cx.nodeOff + slot, 1), len = GA_HN(cx.nodeOff + slot, 3);
					const uint32_t oldEndConf = mem.conf[(size_t)(cs + len - 1) * LANES];
					uint32_t callLastCol = 0;
					const int32_t callMin = ga_ex_calc_node<LANES, SMALL>(g, caps, mem, st, cx, slot, callLastCol);
					if (st.status != GA_OK) return false;
					GA_HN(cx.nodeOff + slot, 2) = (uint32_t)callMin;   // setMinScore: the LAST call's value stays (GraphAligner.h:2375)
					const uint32_t newEndConf = mem.conf[(size_t)(cs + len - 1) * LANES];
					const int32_t endSbs = ga_col_load_sbs<LANES>(mem, cx.slabOff + cs + len - 1);
					if (endSbs < (int32_t)st.partLen && ga_conf_gt((int)(newEndConf & 0xffu), (newEndConf & 0x100u) != 0, (int)(oldEndConf & 0xffu), (oldEndConf & 0x100u) != 0))
					{
						const uint32_t node = GA_HN(cx.nodeOff + slot, 0);
						for (uint32_t e = g.outOff[node], eEnd = g.outOff[node + 1]; e < eEnd; e++)
						{
							int cu = ga_hash_find<LANES>(cx.hashCur, cx.maskCur, cx.stampCur, g.outAdj[e]);
							if (cu < 0 || (mem.cmpOf[(size_t)cu * LANES] & 0x3fffffffu) != ci) continue;
							if ((mem.conf[(size_t)GA_HN(cx.nodeOff + cu, 1) * LANES] & 0xffu) >= 64) continue;
							if (mem.cmpOf[(size_t)cu * LANES] & INQ) continue;
							mem.cmpOf[(size_t)cu * LANES] |= INQ;
							mem.wl[(size_t)(wlN++) * LANES] = (uint32_t)cu;
						}
					}
					if (callMin < minScore || (callMin == minScore && callMin != 0x7fffffff))
					{
						minScore = callMin;
						lastMinSlot = slot;
						lastMinCol = callLastCol;
					}
					if (++guard > 200u * (compEnd - compStart) + 1000u) { st.status = GA_ERR_CYCLE_ITER; return false; }
				}
			}
			compEnd = compStart;
		}
	}
	GA_HDR(cx.s, 10) = lastMinSlot;
	GA_HDR(cx.s, 11) = lastMinCol;
#ifdef GA_HOST_DEBUG
	if (getenv("GA_DBG"))
	{
		std::vector<std::pair<uint32_t, std::pair<int, int>>> v;
		for (uint32_t slot = 0; slot < nc; slot++)
		{
			uint32_t cs = GA_HN(cx.nodeOff + slot, 1), len = GA_HN(cx.nodeOff + slot, 3);
			v.push_back({GA_HN(cx.nodeOff + slot, 0), {(int)GA_HN(cx.nodeOff + slot, 2), ga_tiny_score(GA_TC(cs + len - 1))}});
		}
		std::sort(v.begin(), v.end());
		fprintf(stderr, "SLICE j=%d min=%d n=%d last=%d\n", (int)cx.s * 64, minScore, (int)nc, lastMinSlot == 0xffffffffu ? -2 : (int)(g.nodeStart[GA_HN(cx.nodeOff + lastMinSlot, 0)] + lastMinCol));
		for (auto& p : v) fprintf(stderr, "N %d %d %d\n", (int)p.first, p.second.first, p.second.second);
	}
#endif
	st.wordColumns += ncols;
	// correctness HMM (AlignmentCorrectnessEstimation.cpp:71-89); doubles are only added and compared
	int32_t m = minScore - st.prevMin;
	if (m < 0 || m > 64) { st.status = GA_ERR_INTERNAL; return false; }
	double cc = st.hmmC + hmm.c2c, fc = st.hmmF + hmm.f2c;
	double cf = st.hmmC + hmm.c2f, ff = st.hmmF + hmm.f2f;
	res.correctFromCorrect = cc >= fc;
	res.falseFromCorrect = cf >= ff;
	res.hmmC = (cc > fc ? cc : fc) + hmm.correctMul[m];
	res.hmmF = (cf > ff ? cf : ff) + hmm.falseMul[m];
	res.currentlyCorrect = res.hmmC > res.hmmF;
	res.minScore = minScore;
	return true;
}

template <int LANES>
GA_DEV int ga_slice_find(const GaLaneMem& mem, uint32_t nodeOff, uint32_t nNodes, uint32_t node)
{
	for (uint32_t i = 0; i < nNodes; i++)
	{
		if (GA_HNG(nodeOff + i, 0) == node) return (int)i;
	}
	return -1;
}

// Iterative Tarjan over the band of slice sl in the reference's visiting order (band order, outNeighbors order,
// GraphAligner.h:1759-1856).  Returns the band slot of the first emitted node whose minimum equals minScore; for
// a multi-node component the reference's work-list order decides instead, so there the choice is approximate.
// Scratch: indeg = DFS index (0 = unvisited), order = low link, uorder = Tarjan stack, unext = call stack slots,
// ubkt = call stack edge cursors, tiny[0] bit = on-stack flags are kept in the top bit of indeg.
template <int LANES>
GA_DEV int ga_first_emitted_min_node(const ga_graph_view& g, const ga_caps& caps, const GaLaneMem& mem, int sl, int32_t minScore)
{
	const uint32_t nodeOff = GA_HDR(sl, 2), nNodes = GA_HDR(sl, 3);
	if (nNodes == 1) return 0;
	const uint32_t ONSTACK = 0x80000000u;
	for (uint32_t i = 0; i < nNodes; i++) mem.indeg[(size_t)i * LANES] = 0;
	uint32_t counter = 0;
	uint32_t tstack = 0;
	for (uint32_t root = 0; root < nNodes; root++)
	{
		if (mem.indeg[(size_t)root * LANES] != 0) continue;
		uint32_t depth = 0;
		mem.unext[0] = root;
		mem.ubkt[0] = g.outOff[GA_HNG(nodeOff + root, 0)];
		counter++;
		mem.indeg[(size_t)root * LANES] = counter | ONSTACK;
		mem.order[(size_t)root * LANES] = counter;
		mem.uorder[(size_t)(tstack++) * LANES] = root;
		while (true)
		{
			uint32_t slot = mem.unext[(size_t)depth * LANES];
			uint32_t node = GA_HNG(nodeOff + slot, 0);
			uint32_t e = mem.ubkt[(size_t)depth * LANES];
			if (e < g.outOff[node + 1])
			{
				mem.ubkt[(size_t)depth * LANES] = e + 1;
				int nb = ga_slice_find<LANES>(mem, nodeOff, nNodes, g.outAdj[e]);
				if (nb < 0) continue;
				uint32_t mark = mem.indeg[(size_t)nb * LANES];
				if (mark == 0)
				{
					depth++;
					mem.unext[(size_t)depth * LANES] = (uint32_t)nb;
					mem.ubkt[(size_t)depth * LANES] = g.outOff[GA_HNG(nodeOff + nb, 0)];
					counter++;
					mem.indeg[(size_t)nb * LANES] = counter | ONSTACK;
					mem.order[(size_t)nb * LANES] = counter;
					mem.uorder[(size_t)(tstack++) * LANES] = (uint32_t)nb;
				}
				else if (mark & ONSTACK)
				{
					uint32_t idx = mark & ~ONSTACK;
					if (idx < mem.order[(size_t)slot * LANES]) mem.order[(size_t)slot * LANES] = idx;
				}
				continue;
			}
			// all neighbours done
			uint32_t low = mem.order[(size_t)slot * LANES];
			if (low == (mem.indeg[(size_t)slot * LANES] & ~ONSTACK))
			{
				// emit the component: stack entries down to slot, last pushed first
				int found = -1;
				while (true)
				{
					uint32_t back = mem.uorder[(size_t)(--tstack) * LANES];
					mem.indeg[(size_t)back * LANES] &= ~ONSTACK;
					// within a component the reference evaluates via a LIFO work-list; the root is popped first,
					// so the earliest-listed member is the best guess for "evaluated last"
					if (found < 0 && (int32_t)GA_HNG(nodeOff + back, 2) == minScore) found = (int)back;
					if (back == slot) break;
				}
				if (found >= 0) return found;
			}
			if (depth == 0) break;
			depth--;
			uint32_t parent = mem.unext[(size_t)depth * LANES];
			if (low < mem.order[(size_t)parent * LANES]) mem.order[(size_t)parent * LANES] = low;
		}
	}
	return -1;
}

// ------------------------------------------------------------------------------------------------------------
// End of a stream's forward pass: end trimming (removeWronglyAlignedEnd), the tied minimum cells of the last retained
// slice and the cell the traceback starts from.  Reads the slice headers, node lists and columns from the history in
// global memory; scratch: indeg, order, uorder, unext, ubkt (one entry per band node of a slice).
// ------------------------------------------------------------------------------------------------------------
template <int LANES, bool RAMP = false>
GA_DEV void ga_finish_stream(const ga_graph_view& g, const ga_caps& caps, const GaLaneMem& mem, GaStreamState& st, bool active, uint32_t slicesRun, uint32_t debugFlags, ga_stream_out* out)
{
	// ---- end trimming, trace start, traceback.  No early returns: the traceback is a warp-wide loop --------------------
	bool doTrace = false;
	int n = 0;
	uint32_t endNode = 0, endOff = 0;
	if (active)
	{
		out->nSlicesRun = (int32_t)slicesRun;
		out->wordColumns = st.wordColumns;
		out->cyclicSlices = st.cyclicSlices;
		out->rampRedos = st.rampRedos;
		out->nMapped = 0;
		out->nMoves = 0;
		out->nPathNodes = 0;
		out->nRuns = 0;
		out->nPositions = 0;
		out->nTies = 0;
		out->nSlices = 0;
		out->score = 0;
		out->endNode = 0;
		out->endOff = 0;
		out->status = st.status;
	}
	if (active && st.status == GA_OK)
	{
		// removeWronglyAlignedEnd (GraphAligner.h:2554-2569)
		n = (int)st.slicesPushed;
		if (n > 0)
		{
			bool currentlyCorrect = (GA_HDR(n - 1, 5) & 1u) != 0;
			while (!currentlyCorrect)
			{
				n--;
				if (n == 0) break;
				currentlyCorrect = (GA_HDR(n - 1, 5) & 2u) != 0;   // FalseFromCorrect of the new last slice, as the reference reads it
			}
		}
		out->nSlices = n;
		if (n == 0) out->status = GA_EMPTY;
		else
		{
			// trace start = minScoreIndex.back() of the last retained slice (GraphAligner.h:918-932): the highest tied column
			// of the LAST evaluated node that attains the slice minimum.  Evaluation order = reverse of Tarjan's component
			// emission order over the band (GraphAligner.h:1836-1856,2360), so the wanted node is the first one emitted.
			// (-B ramp: where the reference's last sqrt checkpoint is this very slice it starts from the checkpoint's instance of it)
			const int sl = (RAMP && (GA_HDR(n - 1, 5) & GA_HF_ALT)) ? (int)GA_HDR(n - 1, 1) : n - 1;
			const int32_t minScore = (int32_t)GA_HDR(sl, 4);
			const uint32_t nodeOff = GA_HDR(sl, 2), nNodes = GA_HDR(sl, 3), slabOff = GA_HDR(sl, 0);
			uint32_t nTies = 0;
			for (uint32_t slot = 0; slot < nNodes; slot++)
			{
				if ((int32_t)GA_HNG(nodeOff + slot, 2) != minScore) continue;
				uint32_t node = GA_HNG(nodeOff + slot, 0);
				uint32_t cs = GA_HNG(nodeOff + slot, 1);
				uint32_t len = GA_HNG(nodeOff + slot, 3);
				for (uint32_t k = 0; k < len; k++)
				{
					GaCol c = ga_col_load<LANES>(mem, slabOff + cs + k);
					int32_t v = c.sbs + (int32_t)GA_POPC(c.VP) - (int32_t)GA_POPC(c.VN);
					if (v != minScore) continue;
					if (nTies < GA_MAX_TIES) { out->tieNode[nTies] = node; out->tieOff[nTies] = k; }
					nTies++;
				}
			}
			// a slice with a cyclic component recorded its last minimum cell while replaying the reference's schedule
			const uint32_t recSlot = GA_HDR(sl, 10);
			int endSlot = recSlot != 0xffffffffu ? (int)recSlot : (nTies > 0 ? ga_first_emitted_min_node<LANES>(g, caps, mem, sl, minScore) : -1);
			if (endSlot < 0) out->status = GA_ERR_INTERNAL;
			else
			{
				endNode = GA_HNG(nodeOff + endSlot, 0);
				uint32_t cs = GA_HNG(nodeOff + endSlot, 1);
				uint32_t len = GA_HNG(nodeOff + endSlot, 3);
				if (recSlot != 0xffffffffu) endOff = GA_HDR(sl, 11);
				else
				{
					for (uint32_t k = 0; k < len; k++)
					{
						GaCol c = ga_col_load<LANES>(mem, slabOff + cs + k);
						int32_t v = c.sbs + (int32_t)GA_POPC(c.VP) - (int32_t)GA_POPC(c.VN);
						if (v == minScore) endOff = k;
					}
				}
				out->nTies = nTies;
				out->score = minScore;
				out->endNode = endNode;
				out->endOff = endOff;
				doTrace = !(debugFlags & 1u);
			}
		}
	}
	GA_TLAP(st, 7);
#ifdef GA_PHASE_TIMING
	if (active) for (int i = 0; i < 16; i++) out->phase[i] = st.phase[i];
#endif
	// the traceback kernel (ga_trace.cuh) walks every stream that has a trace start
	if (active) out->traceOff = doTrace ? 1 : 0;
}

// The parts of a read are never materialised: a stream names its source range in the batch's raw read bytes
// (ga_stream_in::seqOff, srcInfo) and the Peq pre-pass reads the characters from there - forwards for the part after the
// seed, backwards and complemented for the part before it (getSplitAlignment, GraphAligner.h:2969-3024 with
// CommonUtils::ReverseComplement), 'N' beyond the real length (the padding to a multiple of 64 rows).
// Character i of the part as (IUPAC match mask, exact code): the mask of a complemented character is the mask with A<->T and
// C<->G exchanged (bit reversal of the 4-bit set); the exact code (GraphAligner.h:1540 compares characters, so only the
// upper-case letters the graph holds can be equal) of a complemented character is always that of an upper-case letter.
GA_DEV void ga_part_char(const uint8_t* raw, const ga_stream_in& in, uint32_t i, uint32_t& mask, uint32_t& code)
{
	const uint32_t real = GA_SRC_LEN(in.srcInfo);
	if (i >= real) { mask = 15; code = 4; return; }
	if (in.srcInfo & GA_SRC_BACKWARD)
	{
		const uint8_t c = raw[in.seqOff - i];
		const uint32_t m = ga_iupac_mask(c);
		mask = ((m & 1u) << 3) | ((m & 2u) << 1) | ((m & 4u) >> 1) | ((m & 8u) >> 3);
		code = mask == 1 ? 0u : (mask == 2 ? 1u : (mask == 4 ? 2u : (mask == 8 ? 3u : 4u)));
	}
	else
	{
		const uint8_t c = raw[in.seqOff + i];
		mask = ga_iupac_mask(c);
		code = ga_exact_code(c);
	}
}

// Match masks of the 64 read characters of slice sl: bit i of {A,C,G,T} word = characterMatch(read[i], base) with IUPAC
// codes (GraphAligner.h:2338-2351).  Used by the Peq pre-pass kernel.
GA_DEV void ga_peq_words(const uint8_t* raw, const ga_stream_in& in, uint32_t sl, uint64_t& BA, uint64_t& BC, uint64_t& BG, uint64_t& BT)
{
	BA = BC = BG = BT = 0;
	for (uint32_t i = 0; i < 64; i++)
	{
		uint32_t m, code;
		ga_part_char(raw, in, sl * 64 + i, m, code);
		BA |= (uint64_t)(m & 1u) << i;
		BC |= (uint64_t)((m >> 1) & 1u) << i;
		BG |= (uint64_t)((m >> 2) & 1u) << i;
		BT |= (uint64_t)((m >> 3) & 1u) << i;
	}
}

// the per-slice word of the Peq pre-pass next to the match masks (GaStreamState::aux)
GA_DEV uint32_t ga_peq_aux(const uint8_t* raw, const ga_stream_in& in, uint32_t sl)
{
	uint32_t m0, c0, mp = 0, cp = 4;
	ga_part_char(raw, in, 0, m0, c0);
	if (sl > 0) ga_part_char(raw, in, sl * 64 - 1, mp, cp);
	return cp | (m0 << 4);
}

// ------------------------------------------------------------------------------------------------------------
// Whole stream, forward part: slices (lock step across the warp), end trimming, tie list, trace start.
// `active` = this lane holds a stream.  warpColTop is the warp-uniform bump pointer into the column slab.
// ------------------------------------------------------------------------------------------------------------
// RAMP (-B > -b, general layout only): the stream also keeps the reference's sqrt checkpoints.  The reference traces back
// through slices it RE-COMPUTES from those checkpoints (getSlicesFromTable, GraphAligner.h:2858-2943), and after a ramp redo
// its pending checkpoint (storeSlice) can be a slice of the abandoned narrow-band pass: the stretch behind such a checkpoint
// is then re-computed from the abandoned slice's end state and differs from what the forward pass found.  To return what the
// reference returns, the checkpoint stack is replayed here (same pushes and pops), and after the forward pass the stretch
// behind every checkpoint that is not the final instance of its slice is computed again from that checkpoint (second phase of
// the slice loop); the traceback (ga_trace.cuh, ALT) then sees each slice as the reference does: its re-computed instance
// when it walks through it, the checkpoint's instance when it looks up from the slice below.
template <int LANES, bool SMALL, bool RAMP = false>
GA_DEV void ga_run_stream(const ga_graph_view& g, const ga_caps& caps, const GaHmmTables& hmm, const GaUmapSchedule& sch, const GaLaneMem& mem, bool active,
	const ga_stream_in* in, const uint32_t* peqAux, int initialBandwidth, int rampBandwidth, uint32_t debugFlags, ga_stream_out* out)
{
	GaStreamState st;
	st.status = GA_OK;
	st.done = !active;
	st.aux = active ? peqAux : nullptr;
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

	// initial slice (getInitialSliceOnlyOneNode, GraphAligner.h:2945-2960): the seed node, every column 0.
	// It lives at node-history entry 0 and in tiny table 0 / hash table 0 with stamp 1.
	uint32_t pNodeOff = 0, pNodes = 0;
	if (!st.done)
	{
		uint32_t len = (uint32_t)(g.nodeStart[st.startNode + 1] - g.nodeStart[st.startNode]);
		if (len > caps.maxCols) { st.status = GA_ERR_COL_OVERFLOW; st.done = true; }
		else
		{
			GA_HN(0, 0) = st.startNode;
			GA_HN(0, 1) = 0;
			GA_HN(0, 2) = 0;
			GA_HN(0, 3) = len;
			GA_HN(0, 4) = g.nodeRec[st.startNode].seqChunk;
			if (SMALL) { GA_HNG(0, 0) = st.startNode; GA_HNG(0, 1) = 0; GA_HNG(0, 2) = 0; GA_HNG(0, 3) = len; GA_HNG(0, 4) = GA_HN(0, 4); }
			ga_hash_insert<LANES>(mem.hash[0], ga_hash_window(1, caps.hashSize), 1, st.startNode, 0);
			for (uint32_t k = 0; k < len; k++) ga_tiny_st<LANES, SMALL>(mem.tiny[0], k, 0);
			pNodes = 1;
			st.histNodeTop = 1;
		}
	}
	// per-lane slice cursor and ramp state (GraphAligner.h:2602-2719); the warp iterates until every lane is done
	static_assert(!(RAMP && SMALL), "the checkpoint replay uses the general layout");
	int ls = 0;
	int rampUntil = 0, rampRedoIndex = -1;
	uint32_t gen = 1;          // stamp generator for the node -> slot tables (stamp 1 = the initial slice in table 0)
	int tp = 0;                // table holding the previous slice
	uint32_t stampPrev = 1;
	uint32_t maskPrev = ga_hash_window(1, caps.hashSize);
	// RAMP: the reference's checkpoint stack (DPTable::slices) and pending checkpoint (storeSlice), GraphAligner.h:2597,2772-2786
	const uint32_t cpBase = caps.maxSlices, cpPending = caps.maxSlices + GA_CP_SLOTS(caps.maxSlices);
	uint32_t cpN = 0;
	int storeS = -1;                 // slice of the pending checkpoint (-1: the initial slice)
	uint32_t storeMem = 28;          // its EstimatedMemoryUsage (GraphAligner.h:136-139): 4 B per column + 28 B per node; initial slice: no cells counted
	uint32_t sampF = 1;              // getSamplingFrequency, GraphAligner.h:2962-2967
	if (RAMP) while ((sampF + 1) * (sampF + 1) <= st.nslices) sampF++;
	int phase = 0;                   // 0 = forward pass, 1 = stretches behind stale checkpoints
	int nKept = 0;                   // slices left by removeWronglyAlignedEnd
	uint32_t rk = 0;                 // phase 1: checkpoint whose stretch [ls, segEnd) is being computed
	int segStart = 0, segEnd = 0;
#define GA_CP_S(k) ((int)GA_HDR(cpBase + (k), 1) - 1)
	// the previous slice becomes the instance with header h again: state from the header, tables rebuilt from its columns
	auto loadPrevState = [&](uint32_t h)
	{
		st.prevMin = (int32_t)GA_HDR(h, 4);
		uint64_t c = (uint64_t)GA_HDR(h, 6) | ((uint64_t)GA_HDR(h, 7) << 32);
		uint64_t f = (uint64_t)GA_HDR(h, 8) | ((uint64_t)GA_HDR(h, 9) << 32);
		st.hmmC = ga_bits_to_double(c);
		st.hmmF = ga_bits_to_double(f);
		pNodeOff = GA_HDR(h, 2);
		pNodes = GA_HDR(h, 3);
	};
	auto rebuildPrevTables = [&](uint32_t h)
	{
		stampPrev = ++gen;
		maskPrev = ga_hash_window(pNodes, caps.hashSize);
		const uint32_t tSlab = GA_HDR(h, 0);
		for (uint32_t i = 0; i < pNodes; i++)
		{
			ga_hash_insert<LANES>(mem.hash[tp], maskPrev, stampPrev, GA_HN(pNodeOff + i, 0), i);
			const uint32_t cs = GA_HN(pNodeOff + i, 1), len = GA_HN(pNodeOff + i, 3);
			for (uint32_t k = 0; k < len; k++)
			{
				GaCol c = ga_col_load<LANES>(mem, tSlab + cs + k);
				ga_tiny_st<LANES, SMALL>(mem.tiny[tp], cs + k, ga_tiny_pack(c, false));
			}
		}
	};
	// phase 1: the next checkpoint from k on that is stale and has slices behind it; false = none left (then every slice whose
	// checkpoint is another instance than the history's gets its GA_HF_ALT reference)
	auto nextStretch = [&](uint32_t k) -> bool
	{
		for (; k < cpN; k++)
		{
			if (!(GA_HDR(cpBase + k, 5) & GA_HF_STALE)) continue;
			const int from = GA_CP_S(k) + 1, to = (k + 1 == cpN) ? nKept : GA_CP_S(k + 1) + 1;
			if (to <= from) continue;
			rk = k; segStart = from; segEnd = to; ls = from;
			loadPrevState(cpBase + k);
			rebuildPrevTables(cpBase + k);
			return true;
		}
		for (k = 1; k < cpN; k++)
		{
			const int cs = GA_CP_S(k);
			if (GA_HDR(cpBase + k, 0) != GA_HDR(cs, 0)) { GA_HDR(cs, 5) |= GA_HF_ALT; GA_HDR(cs, 1) = cpBase + k; }
		}
		return false;
	};
	while (true)
	{
		GA_TLAP(st, 5);
		if (RAMP)
		{
			if (phase == 0 && (st.done || (uint32_t)ls >= st.nslices))
			{
				// the forward pass of this lane is over: removeWronglyAlignedEnd on the checkpoints (GraphAligner.h:2554-2569), then
				// the reference's walk over them (getTraceFromTable, GraphAligner.h:918-941)
				phase = 1;
				st.done = true;
				// (debugFlags bit 1: no replay - the host's second try for a stream whose replayed walk ran into what crashes the reference)
				if (active && st.status == GA_OK && st.rampRedos > 0 && cpN > 0 && !(debugFlags & 2u))
				{
					nKept = (int)st.slicesPushed;
					if (nKept > 0)
					{
						bool currentlyCorrect = (GA_HDR(nKept - 1, 5) & 1u) != 0;
						while (!currentlyCorrect)
						{
							nKept--;
							if (nKept == 0) break;
							currentlyCorrect = (GA_HDR(nKept - 1, 5) & 2u) != 0;
						}
					}
					if (nKept > 0)
					{
						while (cpN > 1 && GA_CP_S(cpN - 1) >= nKept) cpN--;
						// checkpoints out of order: the reference would re-compute an empty stretch and read its first slice (it crashes)
						bool ordered = true, anyStale = false;
						for (uint32_t k = 1; k < cpN; k++)
						{
							if (GA_CP_S(k) <= GA_CP_S(k - 1)) ordered = false;
						}
						// (nothing is re-computed then: the stream keeps the trace of its own forward pass, which the reference does not live to report)
						if (ordered)
						{
							for (uint32_t k = 1; k < cpN; k++)
							{
								const bool stale = GA_HDR(cpBase + k, 0) != GA_HDR(GA_CP_S(k), 0);
								if (stale) { GA_HDR(cpBase + k, 5) |= GA_HF_STALE; anyStale = true; }
							}
							if (anyStale) st.rampRedos |= GA_RAMP_STALE_BIT;
							if (anyStale && nextStretch(1)) st.done = false;
						}
					}
				}
			}
			else if (phase == 1 && !st.done && ls >= segEnd)
			{
				if (!nextStretch(rk + 1)) st.done = true;
			}
		}
		bool run = !st.done && ((RAMP && phase == 1) || (uint32_t)ls < st.nslices);
		if (!GA_WARP_ANY(run)) break;
		const int s = ls;
		const int tc = tp ^ 1;
		uint32_t stampCur = 0;
		uint32_t maskCur = ga_hash_window(pNodes, caps.hashSize);   // bands change slowly: sized from the previous one
		uint32_t ncols = 0;
		int nc = 0;
		uint32_t nodeOff = st.histNodeTop;
		int bandwidth = initialBandwidth;
		if (run)
		{
			if (gen >= 0xfffeu) { st.status = GA_ERR_HIST_OVERFLOW; st.done = true; run = false; }
			else
			{
				stampCur = ++gen;
				// slices up to rampUntil run with rampBandwidth; rampUntil starts at 0, so slice 0 always does (GraphAligner.h:2612)
				bandwidth = (rampUntil >= s) ? rampBandwidth : initialBandwidth;
				// a re-computed slice runs with the bandwidth its final instance ran with (bandwidthPerSlice, GraphAligner.h:2896)
				if (RAMP && phase == 1) bandwidth = (GA_HDR(s, 5) & GA_HF_RAMPBW) ? rampBandwidth : initialBandwidth;
				nc = ga_select_band<LANES, SMALL>(g, caps, sch, mem, st, bandwidth, pNodeOff, pNodes, mem.tiny[tp], mem.hash[tp], maskPrev, stampPrev, mem.hash[tc], maskCur, stampCur, gen, nodeOff, ncols);
				if (nc <= 0) { if (st.status == GA_OK) st.status = GA_ERR_INTERNAL; st.done = true; run = false; ncols = 0; }
			}
		}
		GA_TLAP(st, 0);
		// this slice's columns for all lanes of the warp: one chunk of the global history pool
		uint32_t maxc = GA_WARP_MAX(ncols);
		uint64_t slabOff = GA_POOL_ALLOC(mem.colPoolTop, maxc);
		if (slabOff + maxc > caps.warpCols)
		{
			if (run) { st.status = GA_ERR_COL_OVERFLOW; st.done = true; }
			break;   // warp-uniform
		}
		GA_TLAP(st, 14);
		if (!run) continue;
		slicesRun++;
		GaSliceCtx cx;
		cx.s = s;
		cx.nodeOff = nodeOff;
		cx.nNodes = (uint32_t)nc;
		cx.pNodeOff = pNodeOff;
		cx.pNodes = pNodes;
		cx.slabOff = (uint32_t)slabOff;
		cx.hasPrevSlab = s > 0;
		cx.pSlabOff = s > 0 ? GA_HDR((RAMP && phase == 1 && s == segStart) ? cpBase + rk : (uint32_t)(s - 1), 0) : 0;
		cx.tinyCur = mem.tiny[tc];
		cx.tinyPrev = mem.tiny[tp];
		cx.tinyRef = st.prevMin;
		cx.hashCur = mem.hash[tc];
		cx.hashPrev = mem.hash[tp];
		cx.stampCur = stampCur;
		cx.stampPrev = stampPrev;
		cx.maskCur = maskCur;
		cx.maskPrev = maskPrev;
		GA_HDR(s, 0) = (uint32_t)slabOff;
		GA_HDR(s, 1) = ncols;
		GA_HDR(s, 2) = nodeOff;
		GA_HDR(s, 3) = (uint32_t)nc;
		GaSliceResult res;
		if (!ga_fill_slice<LANES, SMALL>(g, caps, hmm, mem, st, cx, ncols, res)) { st.done = true; continue; }
		if (RAMP && phase == 1)
		{
			// a re-computed slice replaces the history's instance; the stop and ramp rules do not apply (GraphAligner.h:2893-2935)
			st.hmmC = res.hmmC;
			st.hmmF = res.hmmF;
			st.prevMin = res.minScore;
			GA_HDR(s, 4) = (uint32_t)res.minScore;
			pNodeOff = nodeOff;
			pNodes = (uint32_t)nc;
			st.histNodeTop = nodeOff + (uint32_t)nc;
			tp = tc;
			stampPrev = stampCur;
			maskPrev = maskCur;
			ls = s + 1;
			continue;
		}
		// remember where a ramp would restart from (GraphAligner.h:2630-2634)
		if (rampUntil == s - 1 || (rampUntil < s && res.currentlyCorrect && res.falseFromCorrect)) rampRedoIndex = s - 1;
		if (!res.correctFromCorrect) { st.done = true; continue; }   // GraphAligner.h:2640-2647: stop, slice not recorded
		if (!res.currentlyCorrect && rampUntil < s && rampBandwidth > initialBandwidth)
		{
			// ramp: redo from the remembered slice with the wide band up to here (GraphAligner.h:2648-2719)
			const int target = rampRedoIndex;
			rampUntil = s;
			rampRedoIndex = s;
			st.rampRedos++;
			if (target < 0) { st.status = GA_ERR_INTERNAL; st.done = true; continue; }
			// the previous slice becomes slice `target` again: state from its header, tables rebuilt from the history
			loadPrevState((uint32_t)target);
			st.histNodeTop = nodeOff + (uint32_t)nc;   // abandoned entries are simply left behind
			if (SMALL)
			{
				// the ring holds the target's list again; the next list must land right behind it in the ring
				for (uint32_t i = 0; i < pNodes; i++) for (uint32_t f = 0; f < GA_HN_WORDS; f++) GA_HN(pNodeOff + i, f) = GA_HNG(pNodeOff + i, f);
				st.histNodeTop += (pNodeOff + pNodes - st.histNodeTop) & (GA_HN_RING - 1u);
			}
			st.slicesPushed = (uint32_t)target + 1;
			rebuildPrevTables((uint32_t)target);
			// the checkpoints behind the restart point go, the pending one stays (GraphAligner.h:2667)
			if (RAMP) while (cpN > 1 && GA_CP_S(cpN - 1) > target) cpN--;
			ls = target + 1;
			continue;
		}
		// the slice is kept
		if (SMALL)
		{
			for (uint32_t i = 0; i < (uint32_t)nc; i++) for (uint32_t f = 0; f < GA_HN_WORDS; f++) GA_HNG(nodeOff + i, f) = GA_HN(nodeOff + i, f);
		}
		st.hmmC = res.hmmC;
		st.hmmF = res.hmmF;
		st.prevMin = res.minScore;
		GA_HDR(s, 4) = (uint32_t)res.minScore;
		GA_HDR(s, 5) = (res.currentlyCorrect ? 1u : 0u) | (res.falseFromCorrect ? 2u : 0u) | ((RAMP && bandwidth == rampBandwidth) ? GA_HF_RAMPBW : 0u);
		{
			uint64_t c = ga_double_to_bits(res.hmmC), f = ga_double_to_bits(res.hmmF);
			GA_HDR(s, 6) = (uint32_t)c; GA_HDR(s, 7) = (uint32_t)(c >> 32);
			GA_HDR(s, 8) = (uint32_t)f; GA_HDR(s, 9) = (uint32_t)(f >> 32);
		}
		st.slicesPushed = (uint32_t)s + 1;
		if (RAMP)
		{
			// checkpoints (GraphAligner.h:2772-2786): at a sampling point the pending one is pushed unless it is the slice on top of the
			// stack, and this slice becomes pending; any slice cheaper to keep than the pending one replaces it
			const uint32_t memUse = 4u * ncols + 28u * (uint32_t)nc;
			bool capture = false;
			if ((uint32_t)s % sampF == 0 && (cpN == 0 || storeS != GA_CP_S(cpN - 1)))
			{
				if (cpN >= GA_CP_SLOTS(caps.maxSlices)) { st.status = GA_ERR_HIST_OVERFLOW; st.done = true; continue; }
				if (storeS >= 0) for (uint32_t f = 0; f < GA_HDR_WORDS; f++) GA_HDR(cpBase + cpN, f) = GA_HDR(cpPending, f);
				else for (uint32_t f = 0; f < GA_HDR_WORDS; f++) GA_HDR(cpBase + cpN, f) = 0;
				cpN++;
				capture = true;
			}
			if (capture || memUse < storeMem)
			{
				for (uint32_t f = 0; f < GA_HDR_WORDS; f++) GA_HDR(cpPending, f) = GA_HDR(s, f);
				GA_HDR(cpPending, 1) = (uint32_t)s + 1;
				GA_HDR(cpPending, 5) &= ~(GA_HF_STALE | GA_HF_ALT);
				storeS = s;
				storeMem = memUse;
			}
		}
		pNodeOff = nodeOff;
		pNodes = (uint32_t)nc;
		st.histNodeTop = nodeOff + (uint32_t)nc;
		tp = tc;
		stampPrev = stampCur;
		maskPrev = maskCur;
		ls = s + 1;
	}
#undef GA_CP_S
	ga_finish_stream<LANES, RAMP>(g, caps, mem, st, active, slicesRun, debugFlags, out);
}

#endif
