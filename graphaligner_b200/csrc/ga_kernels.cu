// CUDA translation unit: the alignment kernel (one thread per DP stream, one warp per 32 streams) and the
// device context that feeds it.  sm_100a only; there is no CPU fallback - every entry point fails loudly when
// no CUDA device is usable.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <mutex>
#include <chrono>
#include <vector>
#include "ga_core.cuh"
#include "ga_trace.cuh"
#include "ga_fast.cuh"
#include "ga_device.h"

namespace ga
{

#define GA_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " #call); } } while (0)


struct WarpDesc
{
	uint64_t hdrBase;      // element offsets into the pooled arrays (already x LANES)
	uint64_t hnBase;
	uint64_t movesBase;
	uint64_t pathBase;
	uint64_t runsBase;
	uint32_t maxSlices;
	uint32_t histNodes;
	uint32_t maxMoves;
	uint32_t maxPathNodes;
	uint32_t maxRuns;
};

struct ScratchPtrs
{
	uint32_t* tiny;     // [warp][3][maxCols][S]  tiny x2, confirmedRows (cyclic slices)
	uint64_t* hash;     // [warp][2][hashSize][S]
	uint64_t* heap;     // [warp][maxQueue][S]
	uint32_t* nodeTmp;  // [warp][10][maxNodes][S]  indeg, order, unext, uorder, nWlo, nWhi, nPcs, cmpOf, emit, wl
	uint32_t* ubkt;     // [warp][ubktSize][S]
	uint32_t* hdr;
	uint32_t* histNode;
	uint4* colVV;       // column history pool: {VP, VN} [column][S]
	uint32_t* colS;     //   row -1 score | flags [column][S]
	uint32_t* moves;
	uint32_t* pathNodes;
	uint32_t* runs;
	const uint4* peq;   // [stream][slice][2]
	const uint64_t* peqOff;  // per stream: first uint4 of its masks
	uint32_t* peqAux;        // per stream and slice (index peqOff / 2 + slice): exact code of the read character above the slice | IUPAC mask of the first character << 4
	unsigned long long* colPoolTop;
	uint64_t colPoolCap;
	uint32_t ubktSize;
};

__constant__ GaHmmTables c_hmm;
__constant__ GaUmapSchedule c_sched;

#ifndef GA_HOSTSIM
// Character table of the two pre-pass kernels in shared memory: IUPAC match mask (4 bits, 0 = a character the reference
// aborts on) | exact code << 4, from the same two functions the rest of the code uses (ga_iupac_mask, ga_exact_code)
__device__ __forceinline__ void ga_fill_char_table(uint8_t* tab)
{
	for (uint32_t c = threadIdx.x; c < 256; c += blockDim.x) tab[c] = (uint8_t)(ga_iupac_mask((uint8_t)c) | (ga_exact_code((uint8_t)c) << 4));
	__syncthreads();
}

// character i of a stream's part as ga_part_char gives it (mask, exact code), through the table
__device__ __forceinline__ void ga_part_char_tab(const uint8_t* __restrict__ raw, const uint8_t* tab, const ga_stream_in& in, uint32_t real, uint32_t i, uint32_t& mask, uint32_t& code)
{
	if (i >= real) { mask = 15; code = 4; return; }
	if (in.srcInfo & GA_SRC_BACKWARD)
	{
		const uint32_t m = tab[raw[in.seqOff - i]] & 15u;
		mask = ((m & 1u) << 3) | ((m & 2u) << 1) | ((m & 4u) >> 1) | ((m & 8u) >> 3);
		code = mask == 1 ? 0u : (mask == 2 ? 1u : (mask == 4 ? 2u : (mask == 8 ? 3u : 4u)));
	}
	else
	{
		const uint32_t t = tab[raw[in.seqOff + i]];
		mask = t & 15u;
		code = t >> 4;
	}
}

// Match masks for every 64-row slice of every stream (ga_peq_words / ga_peq_aux): one block per stream, one WARP per slice -
// lane l reads characters l and l + 32 of the slice (two coalesced 32-byte rows of the read), the four match words are the
// ballots of the characters' mask bits.  64 bytes in, 36 bytes out per slice.
__global__ void __launch_bounds__(128) ga_peq_kernel(const ga_stream_in* __restrict__ streams, const uint64_t* __restrict__ peqOff, const uint8_t* __restrict__ parts, uint32_t nStreams, uint4* __restrict__ peq,
	uint32_t* __restrict__ peqAux)
{
	__shared__ uint8_t tab[256];
	ga_fill_char_table(tab);
	const uint32_t stream = blockIdx.x;
	if (stream >= nStreams) return;
	const ga_stream_in in = streams[stream];
	const uint32_t real = GA_SRC_LEN(in.srcInfo);
	const uint32_t nslices = in.partLen / 64;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	const uint64_t off = peqOff[stream];
	uint32_t m0, c0;
	ga_part_char_tab(parts, tab, in, real, 0, m0, c0);
	for (uint32_t sl = warp; sl < nslices; sl += nwarps)
	{
		uint32_t ma, ca, mb, cb;
		ga_part_char_tab(parts, tab, in, real, sl * 64 + lane, ma, ca);
		ga_part_char_tab(parts, tab, in, real, sl * 64 + 32 + lane, mb, cb);
		uint32_t w[8];
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			w[2 * k] = __ballot_sync(0xffffffffu, (ma >> k) & 1u);
			w[2 * k + 1] = __ballot_sync(0xffffffffu, (mb >> k) & 1u);
		}
		// the exact code of the character above the slice = the last character of the slice before (lane 31's second character)
		const uint32_t prevCode = __shfl_sync(0xffffffffu, cb, 31);
		if (lane == 0)
		{
			uint4* dst = peq + off + (size_t)sl * 2;
			dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
			dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
			if (sl + 1 < nslices) peqAux[off / 2 + sl + 1] = prevCode | (m0 << 4);
			if (sl == 0) peqAux[off / 2] = 4u | (m0 << 4);
		}
	}
}

// Reads the reference aborts on (a character outside its IUPAC switch, GraphAligner.h:2039-2110): one block per read, four
// characters per thread and pass
__global__ void __launch_bounds__(128) ga_validate_kernel(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ readOff, uint32_t nReads, uint32_t* __restrict__ bad)
{
	__shared__ uint8_t tab[256];
	ga_fill_char_table(tab);
	const uint32_t read = blockIdx.x;
	if (read >= nReads) return;
	const uint8_t* p = raw + readOff[read];
	const uint64_t len = readOff[read + 1] - readOff[read];
	uint32_t ok = 15u;
	uint64_t i = threadIdx.x;
	for (; i + 3 * blockDim.x < len; i += 4 * blockDim.x)
	{
		const uint8_t a = p[i], b = p[i + blockDim.x], c = p[i + 2 * blockDim.x], d = p[i + 3 * blockDim.x];
		ok = (tab[a] & 15u) && (tab[b] & 15u) && (tab[c] & 15u) && (tab[d] & 15u) ? ok : 0u;
	}
	for (; i < len; i += blockDim.x) ok = (tab[p[i]] & 15u) ? ok : 0u;
	const int any = __syncthreads_or(ok == 0u ? 1 : 0);
	if (threadIdx.x == 0) bad[read] = any ? 1u : 0u;
}
#endif

// shared-memory scratch per stream in small-band mode
#define GA_SMEM_NODES 16u
#define GA_SMEM_HASH 32u
#define GA_SMEM_HEAP 32u
#define GA_SMEM_UBKT 32u
#define GA_SMEM_COLS 256u
#define GA_SMEM_WORDS64 (2 * GA_SMEM_HASH + GA_SMEM_HEAP + 2 * GA_SMEM_NODES + (8 * GA_SMEM_NODES + GA_SMEM_UBKT + GA_HN_RING * GA_HN_WORDS) / 2 + (2 * GA_SMEM_COLS) / 4)

// Per-lane pointers of one stream (lane ml of `warp`, S streams per warp).  In small-band mode (host: caps.maxNodes =
// GA_SMEM_NODES, hashSize = GA_SMEM_HASH, maxQueue = GA_SMEM_HEAP, maxCols = GA_SMEM_COLS) everything that is written and
// read back within a slice or by the next one lives in the warp's shared-memory block `ws` (GA_SMEM_WORDS64 x S words):
// a global store invalidates its L1 line, so in global memory every such read-after-write is an L2 round trip on the
// stream's critical path.  eqTab = the warp's four match words, [base][lane].
static __host__ __device__ inline void setupLaneMem(GaLaneMem& mem, const ScratchPtrs& sp, const WarpDesc& wd, const ga_caps& caps, size_t w, uint32_t ml, uint32_t S, bool small,
	unsigned long long* ws, uint64_t* eqTab)
{
	mem.conf = sp.tiny + (w * 3 + 2) * caps.maxCols * S + ml;
	mem.cmpOf = sp.nodeTmp + (w * 10 + 7) * caps.maxNodes * S + ml;
	mem.emit = sp.nodeTmp + (w * 10 + 8) * caps.maxNodes * S + ml;
	mem.wl = sp.nodeTmp + (w * 10 + 9) * caps.maxNodes * S + ml;
	mem.hdr = sp.hdr + wd.hdrBase + ml;
	mem.histNode = sp.histNode + wd.hnBase + ml;
	mem.colVV = sp.colVV + ml;
	mem.colS = sp.colS + ml;
	mem.colPoolTop = sp.colPoolTop;
	mem.eqTab = eqTab + ml;
	mem.hnRing = nullptr;
	mem.lastVV = nullptr;
	mem.lastS = nullptr;
	if (!small)
	{
		mem.tiny[0] = sp.tiny + (w * 3 + 0) * caps.maxCols * S + ml;
		mem.tiny[1] = sp.tiny + (w * 3 + 1) * caps.maxCols * S + ml;
		mem.hash[0] = sp.hash + (w * 2 + 0) * caps.hashSize * S + ml;
		mem.hash[1] = sp.hash + (w * 2 + 1) * caps.hashSize * S + ml;
		mem.heap = sp.heap + w * caps.maxQueue * S + ml;
		mem.indeg = sp.nodeTmp + (w * 10 + 0) * caps.maxNodes * S + ml;
		mem.order = sp.nodeTmp + (w * 10 + 1) * caps.maxNodes * S + ml;
		mem.unext = sp.nodeTmp + (w * 10 + 2) * caps.maxNodes * S + ml;
		mem.uorder = sp.nodeTmp + (w * 10 + 3) * caps.maxNodes * S + ml;
		mem.nWlo = sp.nodeTmp + (w * 10 + 4) * caps.maxNodes * S + ml;
		mem.nWhi = sp.nodeTmp + (w * 10 + 5) * caps.maxNodes * S + ml;
		mem.nPcs = sp.nodeTmp + (w * 10 + 6) * caps.maxNodes * S + ml;
		mem.ubkt = sp.ubkt + w * sp.ubktSize * S + ml;
		return;
	}
	mem.hash[0] = (uint64_t*)ws + ml;
	mem.hash[1] = (uint64_t*)ws + GA_SMEM_HASH * S + ml;
	mem.heap = (uint64_t*)ws + 2 * GA_SMEM_HASH * S + ml;
	mem.lastVV = (uint64_t*)ws + (2 * GA_SMEM_HASH + GA_SMEM_HEAP) * S + ml;
	uint32_t* base32 = (uint32_t*)(ws + (2 * GA_SMEM_HASH + GA_SMEM_HEAP + 2 * GA_SMEM_NODES) * S);
	mem.indeg = base32 + 0 * GA_SMEM_NODES * S + ml;
	mem.order = base32 + 1 * GA_SMEM_NODES * S + ml;
	mem.unext = base32 + 2 * GA_SMEM_NODES * S + ml;
	mem.uorder = base32 + 3 * GA_SMEM_NODES * S + ml;
	mem.nWlo = base32 + 4 * GA_SMEM_NODES * S + ml;
	mem.nWhi = base32 + 5 * GA_SMEM_NODES * S + ml;
	mem.nPcs = base32 + 6 * GA_SMEM_NODES * S + ml;
	mem.lastS = base32 + 7 * GA_SMEM_NODES * S + ml;
	mem.ubkt = base32 + 8 * GA_SMEM_NODES * S + ml;
	mem.hnRing = base32 + (8 * GA_SMEM_NODES + GA_SMEM_UBKT) * S + ml;
	uint16_t* base16 = (uint16_t*)(base32 + (8 * GA_SMEM_NODES + GA_SMEM_UBKT + GA_HN_RING * GA_HN_WORDS) * S);
	mem.tiny[0] = base16 + ml;
	mem.tiny[1] = base16 + GA_SMEM_COLS * S + ml;
}

// S = streams per warp (lanes S..31 idle).  Small batches run with small S: more warps to hide latency and
// less divergence; big batches run with S = 32 for full lane utilisation.
#ifndef GA_HOSTSIM
template <int S, bool SMALL, bool RAMP = false>
__global__ void __launch_bounds__(64, 10) ga_forward_kernel(ga_graph_view g, ga_caps caps, ScratchPtrs sp, const WarpDesc* __restrict__ warpDescs,
	const ga_stream_in* __restrict__ streams, uint32_t nStreams, int initialBandwidth, int rampBandwidth, uint32_t debugFlags,
	ga_stream_out* __restrict__ outs)
{
	extern __shared__ __align__(16) unsigned long long gaShared[];
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t warp = tid >> 5;
	const uint32_t lane = tid & 31;
	if ((uint64_t)warp * S >= nStreams) return;   // whole warp out of range (warp-uniform)
	const uint32_t stream = warp * S + lane;
	const bool active = lane < S && stream < nStreams;
	const uint32_t ml = lane < S ? lane : 0;      // idle lanes alias lane 0's addresses but never touch memory
	const WarpDesc wd = warpDescs[warp];
	ga_caps wc = caps;
	wc.maxSlices = wd.maxSlices;
	wc.histNodes = wd.histNodes;
	wc.warpCols = sp.colPoolCap;
	GaLaneMem mem;
	// shared memory: four match words per stream [warp in block][base][lane], then (small-band mode) the warps' scratch blocks
	uint64_t* eqTab = (uint64_t*)gaShared + (size_t)(threadIdx.x >> 5) * (4 * S);
	unsigned long long* ws = gaShared + (size_t)(blockDim.x >> 5) * (4 * S) + (size_t)(threadIdx.x >> 5) * (GA_SMEM_WORDS64 * S);
	if (SMALL)
	{
		for (uint32_t i = lane; i < (2 * GA_SMEM_HASH) * S; i += 32) ws[i] = 0;   // the stamped tables start empty
		__syncwarp();
	}
	setupLaneMem(mem, sp, wd, caps, warp, ml, S, SMALL, ws, eqTab);
	mem.peq = active ? sp.peq + sp.peqOff[stream] : nullptr;
	ga_stream_out* out = active ? outs + stream : nullptr;
	ga_run_stream<S, SMALL, RAMP>(g, wc, c_hmm, c_sched, mem, active, active ? streams + stream : nullptr, active ? sp.peqAux + sp.peqOff[stream] / 2 : nullptr, initialBandwidth, rampBandwidth, debugFlags, out);
}
#endif

#ifndef GA_HOSTSIM
// Small-band forward kernel (ga_fast.cuh): one warp per block, S streams per warp, the warp's state in its block's shared memory
// (32, 1): shared memory admits four of these blocks per SM at most, so ptxas may take the registers it wants (119, no spills;
// with its default budget 96 and 24 bytes of spills - 8.19 vs 8.10 ms, within the noise)
template <int S>
__global__ void __launch_bounds__(32, 1) ga_fast_kernel(ga_graph_view g, ga_caps caps, ScratchPtrs sp, const WarpDesc* __restrict__ warpDescs,
	const ga_stream_in* __restrict__ streams, uint32_t nStreams, int initialBandwidth, int rampBandwidth, uint32_t debugFlags, ga_stream_out* __restrict__ outs)
{
	extern __shared__ __align__(16) unsigned long long gaShared[];
	GaFastShared<S>& sh = *reinterpret_cast<GaFastShared<S>*>(gaShared);
	const uint32_t warp = blockIdx.x;
	const uint32_t lane = threadIdx.x;
	const uint32_t stream = warp * S + lane;
	const bool active = lane < S && stream < nStreams;
	const uint32_t ml = lane < S ? lane : 0;      // idle lanes alias lane 0's addresses but never touch memory
	const WarpDesc wd = warpDescs[warp];
	ga_caps wc = caps;
	wc.maxSlices = wd.maxSlices;
	wc.histNodes = wd.histNodes;
	wc.warpCols = sp.colPoolCap;
	const GaFastLane<S> fl(sh, ml);
	GaLaneMem mem;
	memset(&mem, 0, sizeof(mem));
	mem.hdr = sp.hdr + wd.hdrBase + ml;
	mem.histNode = sp.histNode + wd.hnBase + ml;
	mem.colVV = sp.colVV + ml;
	mem.colS = sp.colS + ml;
	mem.colPoolTop = sp.colPoolTop;
	mem.peq = active ? sp.peq + sp.peqOff[stream] : nullptr;
	// scratch of ga_finish_stream (Tarjan over the last slice's band): the selection scratch and the queue
	mem.indeg = fl.scratch();
	mem.order = fl.scratch() + 16 * S;
	mem.uorder = fl.scratch() + 32 * S;
	mem.unext = fl.scratch() + 48 * S;
	mem.ubkt = (uint32_t*)&sh.heap[0][0] + ml;
	ga_fast_stream<S>(g, wc, c_hmm, c_sched, mem, fl, active, active ? streams + stream : nullptr, active ? sp.peqAux + sp.peqOff[stream] / 2 : nullptr, initialBandwidth, rampBandwidth,
		debugFlags, active ? outs + stream : nullptr);
}
#endif

// Pointers of one stream's trace inputs and temporary outputs (the forward launch's layout: [..][S] interleaved per warp)
static __host__ __device__ inline GaTraceMem traceMemOf(const ScratchPtrs& sp, const WarpDesc& wd, uint32_t stream, uint32_t S)
{
	const uint32_t fl = stream % S;
	GaTraceMem tm;
	tm.S = S;
	tm.hdr = sp.hdr + wd.hdrBase + fl;
	tm.histNode = sp.histNode + wd.hnBase + fl;
	tm.colVV = sp.colVV + fl;
	tm.colS = sp.colS + fl;
	tm.peq = sp.peq + sp.peqOff[stream];
	tm.moves = sp.moves + wd.movesBase + fl;
	tm.pathNodes = sp.pathNodes + wd.pathBase + fl;
	tm.runs = sp.runs + wd.runsBase + fl;
	tm.maxMoves = wd.maxMoves;
	tm.maxPathNodes = wd.maxPathNodes;
	tm.maxRuns = wd.maxRuns;
	return tm;
}

#ifndef GA_HOSTSIM
// Traceback (ga_trace.cuh): one warp per block walks T streams (lane per stream, all 32 lanes fetch the windows), the
// windows and slice tables in the block's shared memory.  Runs after the forward kernel on the same stream.
#ifndef GA_TRACE_MINBLOCKS
#define GA_TRACE_MINBLOCKS 12   /* blocks per SM the traceback kernel's register budget must allow: 12 = 170 registers (it uses 168), one wave of 1667 six-stream warps on 148 SMs; measured: 16 blocks (128 registers) 6.1 ms, 14 (146) 6.2 ms, 12 (168) 5.3 ms at T = 6 */
#endif
template <int T, int P, bool ALT = false>
__global__ void __launch_bounds__(32, GA_TRACE_MINBLOCKS) ga_trace_kernel(ga_graph_view g, ScratchPtrs sp, const WarpDesc* __restrict__ warpDescs, const ga_stream_in* __restrict__ streams,
	uint32_t nStreams, uint32_t S, ga_stream_out* __restrict__ outs, uint32_t* __restrict__ arena, unsigned long long* arenaTop, unsigned long long arenaCap)
{
	extern __shared__ __align__(16) unsigned long long gaShared[];
	GaTraceShared<T, P>& sh = *reinterpret_cast<GaTraceShared<T, P>*>(gaShared);
	const uint32_t lane = threadIdx.x;
	const uint32_t stream = blockIdx.x * T + lane;
	const bool have = lane < T && stream < nStreams;
	ga_stream_out* out = have ? outs + stream : nullptr;
	const bool doTrace = have && out->traceOff != 0 && out->status == GA_OK;
	GaTraceLane L;
	memset(&L.tm, 0, sizeof(L.tm));
	L.tm.S = S;
	L.in = have ? streams + stream : streams;
	if (have) L.tm = traceMemOf(sp, warpDescs[stream / S], stream, S);
	ga_trace_init(L, doTrace, doTrace ? out->nSlices : 0, doTrace ? out->endNode : 0, doTrace ? out->endOff : 0, doTrace ? out->score : 0);
	ga_trace_warp<T, P, ALT>(g, sh, lane, &L, S);
	// compact the streams' trace records into the arena: the warp copies one stream's record at a time, 32 words per step
	const GaTraceMem& tm = L.tm;
	const uint32_t nMoves = L.t.nMoves, nPath = L.t.nPath, nRuns = L.t.nRuns;
	const uint32_t moveWords = (nMoves + 15) / 16;
	// a stream that is all its read has (GA_SRC_SOLO) leaves the read's final mapping records instead of its runs: the host then
	// neither reads runs nor writes mappings for it (the records are D2H'd straight into the result block)
	// (no run on one of the graph's two dummy nodes - ga_tr_runs_mappable, checked by the 32 lanes together, a stream at a time)
	const bool candidate = have && doTrace && L.t.status == GA_OK && (L.in->srcInfo & GA_SRC_SOLO) != 0 && nRuns > 0;
	uint32_t mappedMask = 0;
	for (int r = 0; r < T; r++)
	{
		if (!__shfl_sync(0xffffffffu, candidate ? 1u : 0u, r)) continue;
		const uint32_t nr = __shfl_sync(0xffffffffu, nRuns, r);
		const uint32_t* pr = (const uint32_t*)__shfl_sync(0xffffffffu, (unsigned long long)tm.runs, r);
		bool dummy = false;
		for (uint32_t j = lane; j < nr; j += 32)
		{
			const uint32_t node = pr[(size_t)(j * GA_RUN_WORDS) * S];
			dummy = dummy || node == 0 || node + 1 == g.nNodes;
		}
		if (!__any_sync(0xffffffffu, dummy)) mappedMask |= 1u << r;
	}
	const bool mapped = candidate && ((mappedMask >> lane) & 1u) != 0;
	const uint32_t runWords = mapped ? nRuns * GA_MAP_WORDS + 7u : nRuns * GA_RUN_WORDS;   // + 7: room to start on a 32-byte boundary
	const uint32_t words = have ? moveWords + nPath + runWords : 0;
	const unsigned long long off = words ? atomicAdd(arenaTop, (unsigned long long)words) : 0ull;
	const bool fits = off + words <= arenaCap;
	if (have)
	{
		out->traceOff = off;
		out->nMapped = 0;
		if (doTrace)
		{
			out->nMoves = nMoves;
			out->nPathNodes = nPath;
			out->nRuns = nRuns;
			out->nMapped = mapped ? nRuns : 0;
			out->nPositions = ga_trace_positions(L.t);
			out->status = !fits && L.t.status == GA_OK ? GA_ERR_TRACE_OVERFLOW : L.t.status;
		}
	}
	for (int r = 0; r < T; r++)
	{
		const uint32_t w = __shfl_sync(0xffffffffu, (fits && words) ? words : 0u, r);
		if (w == 0) continue;
		const uint32_t mw = __shfl_sync(0xffffffffu, moveWords, r), np = __shfl_sync(0xffffffffu, nPath, r), nr = __shfl_sync(0xffffffffu, nRuns, r);
		const bool mp = __shfl_sync(0xffffffffu, mapped ? 1u : 0u, r) != 0;
		const unsigned long long o = __shfl_sync(0xffffffffu, off, r);
		const uint32_t* pm = (const uint32_t*)__shfl_sync(0xffffffffu, (unsigned long long)tm.moves, r);
		const uint32_t* pp = (const uint32_t*)__shfl_sync(0xffffffffu, (unsigned long long)tm.pathNodes, r);
		const uint32_t* pr = (const uint32_t*)__shfl_sync(0xffffffffu, (unsigned long long)tm.runs, r);
		for (uint32_t i = lane; i < mw; i += 32) arena[o + i] = pm[(size_t)i * S];
		for (uint32_t i = lane; i < np; i += 32) arena[o + mw + i] = pp[(size_t)i * S];
		if (mp)
		{
			const unsigned long long mo = o + mw + np + GA_MAP_PAD(o + mw + np);
			GaDeviceMapping* dst = (GaDeviceMapping*)(arena + mo);
			for (uint32_t j = lane; j < nr; j += 32) dst[j] = ga_tr_mapping(g, pr, S, nr, j);
		}
		else for (uint32_t i = lane; i < nr * GA_RUN_WORDS; i += 32) arena[o + mw + np + i] = pr[(size_t)i * S];
	}
}
#endif

#ifndef GA_HOSTSIM
// INT32 roofline probe: 8 independent dependency chains per thread of alternating LOP3 / IADD3, no memory traffic.
// Its rate is the denominator of the integer-ALU roofline fraction.
__global__ void ga_int32_peak_kernel(uint32_t* sink, int iters)
{
	uint32_t a0 = threadIdx.x, a1 = blockIdx.x, a2 = 0x9e3779b9u, a3 = 0x7f4a7c15u, a4 = 1, a5 = 2, a6 = 3, a7 = 4;
	uint32_t k = threadIdx.x * 2654435761u + 1;
#pragma unroll 1
	for (int i = 0; i < iters; i++)
	{
#pragma unroll
		for (int u = 0; u < 8; u++)
		{
			a0 = (a0 ^ k) & (a0 | 0x55555555u); a1 = a1 + k + 3; a2 = (a2 ^ k) & (a2 | 0x33333333u); a3 = a3 + k + 5;
			a4 = (a4 ^ k) & (a4 | 0x0f0f0f0fu); a5 = a5 + k + 7; a6 = (a6 ^ k) & (a6 | 0x00ff00ffu); a7 = a7 + k + 9;
		}
	}
	if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x12345678u) sink[0] = a0;
}
#endif

#ifdef GA_HOSTSIM
// ---- CPU emulation of the launches (test infrastructure, see oracle/hostsim/cuda_runtime.h): one stream per "warp" ----
static void hostsim_peq(const ga_stream_in* streams, const uint64_t* peqOff, const uint8_t* parts, uint32_t nStreams, uint4* peq, uint32_t* peqAux)
{
	for (uint32_t stream = 0; stream < nStreams; stream++)
	{
		const ga_stream_in in = streams[stream];
		const uint32_t nslices = in.partLen / 64;
		for (uint32_t sl = 0; sl < nslices; sl++)
		{
			uint64_t A, C, G, T;
			ga_peq_words(parts, in, sl, A, C, G, T);
			uint4* dst = peq + peqOff[stream] + (size_t)sl * 2;
			dst[0] = make_uint4((uint32_t)A, (uint32_t)(A >> 32), (uint32_t)C, (uint32_t)(C >> 32));
			dst[1] = make_uint4((uint32_t)G, (uint32_t)(G >> 32), (uint32_t)T, (uint32_t)(T >> 32));
			peqAux[peqOff[stream] / 2 + sl] = ga_peq_aux(parts, in, sl);
		}
	}
}

static void hostsim_align(ga_graph_view g, ga_caps caps, ScratchPtrs sp, const WarpDesc* warpDescs, const ga_stream_in* streams, const uint8_t* parts, uint32_t nStreams,
	int initialBandwidth, int rampBandwidth, uint32_t debugFlags, bool small, ga_stream_out* outs, uint32_t* arena, unsigned long long* arenaTop, unsigned long long arenaCap)
{
	const int S = 1;
	std::vector<uint64_t> eqTab(4);
	std::vector<unsigned long long> ws(GA_SMEM_WORDS64);
	for (uint32_t stream = 0; stream < nStreams; stream++)
	{
		const size_t w = stream;
		const WarpDesc wd = warpDescs[w];
		ga_caps wc = caps;
		wc.maxSlices = wd.maxSlices;
		wc.histNodes = wd.histNodes;
		wc.warpCols = sp.colPoolCap;
		GaLaneMem mem;
		if (small)
		{
			// the small-band kernel body with one lane
			static GaFastShared<1> sh;
			const GaFastLane<1> fl(sh, 0);
			memset(&mem, 0, sizeof(mem));
			mem.hdr = sp.hdr + wd.hdrBase;
			mem.histNode = sp.histNode + wd.hnBase;
			mem.colVV = sp.colVV;
			mem.colS = sp.colS;
			mem.colPoolTop = sp.colPoolTop;
			mem.peq = sp.peq + sp.peqOff[stream];
			mem.indeg = fl.scratch();
			mem.order = fl.scratch() + 16;
			mem.uorder = fl.scratch() + 32;
			mem.unext = fl.scratch() + 48;
			mem.ubkt = (uint32_t*)&sh.heap[0][0];
			ga_fast_stream<1>(g, wc, c_hmm, c_sched, mem, fl, true, streams + stream, sp.peqAux + sp.peqOff[stream] / 2, initialBandwidth, rampBandwidth, debugFlags, outs + stream);
			continue;
		}
		std::fill(ws.begin(), ws.end(), 0ull);
		setupLaneMem(mem, sp, wd, caps, w, 0, S, false, ws.data(), eqTab.data());
		mem.peq = sp.peq + sp.peqOff[stream];
		if (rampBandwidth > initialBandwidth) ga_run_stream<1, false, true>(g, wc, c_hmm, c_sched, mem, true, streams + stream, sp.peqAux + sp.peqOff[stream] / 2, initialBandwidth, rampBandwidth, debugFlags, outs + stream);
		else ga_run_stream<1, false>(g, wc, c_hmm, c_sched, mem, true, streams + stream, sp.peqAux + sp.peqOff[stream] / 2, initialBandwidth, rampBandwidth, debugFlags, outs + stream);
	}
	// the traceback "launch": the same warp code with its lanes as loops, HT streams per "warp"
	const int HT = 4, HP = 1;
	for (uint32_t first = 0; first < nStreams; first += HT)
	{
		static GaTraceShared<HT, HP> sh;
		GaTraceLane lanes[HT];
		bool doTrace[HT];
		for (int l = 0; l < HT; l++)
		{
			const uint32_t stream = first + l;
			const bool have = stream < nStreams;
			ga_stream_out* out = have ? outs + stream : nullptr;
			doTrace[l] = have && out->traceOff != 0 && out->status == GA_OK;
			memset(&lanes[l].tm, 0, sizeof(GaTraceMem));
			lanes[l].tm.S = S;
			lanes[l].in = have ? streams + stream : streams;
			if (have) lanes[l].tm = traceMemOf(sp, warpDescs[stream / S], stream, S);
			ga_trace_init(lanes[l], doTrace[l], doTrace[l] ? out->nSlices : 0, doTrace[l] ? out->endNode : 0, doTrace[l] ? out->endOff : 0, doTrace[l] ? out->score : 0);
		}
		if (rampBandwidth > initialBandwidth) ga_trace_warp<HT, HP, true>(g, sh, 0, lanes, S);
		else ga_trace_warp<HT, HP>(g, sh, 0, lanes, S);
		for (int l = 0; l < HT; l++)
		{
			const uint32_t stream = first + l;
			if (stream >= nStreams) break;
			ga_stream_out* out = outs + stream;
			const GaTraceMem& tm = lanes[l].tm;
			const uint32_t nMoves = lanes[l].t.nMoves, nPath = lanes[l].t.nPath, nRuns = lanes[l].t.nRuns;
			const uint32_t moveWords = (nMoves + 15) / 16;
			const bool mapped = doTrace[l] && lanes[l].t.status == GA_OK && (lanes[l].in->srcInfo & GA_SRC_SOLO) != 0 && ga_tr_runs_mappable(tm, nRuns, g.nNodes);
			const uint32_t runWords = mapped ? nRuns * GA_MAP_WORDS + 7u : nRuns * GA_RUN_WORDS;
			const uint32_t words = moveWords + nPath + runWords;
			const unsigned long long off = *arenaTop;
			*arenaTop += words;
			out->traceOff = off;
			out->nMapped = 0;
			if (doTrace[l])
			{
				out->nMoves = nMoves;
				out->nPathNodes = nPath;
				out->nRuns = nRuns;
				out->nMapped = mapped ? nRuns : 0;
				out->nPositions = ga_trace_positions(lanes[l].t);
				out->status = off + words > arenaCap && lanes[l].t.status == GA_OK ? GA_ERR_TRACE_OVERFLOW : lanes[l].t.status;
			}
			if (off + words > arenaCap) continue;
			for (uint32_t i = 0; i < moveWords; i++) arena[off + i] = tm.moves[(size_t)i * S];
			for (uint32_t i = 0; i < nPath; i++) arena[off + moveWords + i] = tm.pathNodes[(size_t)i * S];
			if (mapped)
			{
				const unsigned long long mo = off + moveWords + nPath + GA_MAP_PAD(off + moveWords + nPath);
				GaDeviceMapping* dst = (GaDeviceMapping*)(arena + mo);
				for (uint32_t j = 0; j < nRuns; j++) dst[j] = ga_tr_mapping(g, tm.runs, S, nRuns, j);
			}
			else for (uint32_t i = 0; i < nRuns * GA_RUN_WORDS; i++) arena[off + moveWords + nPath + i] = tm.runs[(size_t)i * S];
		}
	}
}
#endif

// ---- device buffer pool -----------------------------------------------------------------------------------------
struct Buffer
{
	void* ptr = nullptr;
	size_t cap = 0;
	void ensure(size_t bytes)
	{
		if (bytes <= cap) return;
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		cudaError_t e = cudaMalloc(&ptr, want);
		if (e != cudaSuccess)
		{
			ptr = nullptr;
			throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e) + " allocating " + std::to_string(want >> 20) + " MiB of device memory");
		}
		cap = want;
	}
	void release()
	{
		if (ptr) cudaFree(ptr);
		ptr = nullptr;
		cap = 0;
	}
};

struct DeviceCtx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	std::string lastError;
	// graph
	Buffer gNodeStart, gSeq, gInOff, gInAdj, gOutOff, gOutAdj, gNodeRec, gChunks, gNodeIdRev;
	ga_graph_view view;
	size_t graphBytes = 0;
	bool hasGraph = false;
	// batch buffers
	Buffer bParts, bIn, bOut, bWd, bTiny, bHash, bHeap, bNodeTmp, bUbkt, bHdr, bHn, bColVV, bColS, bPeq, bPeqAux, bPeqOff, bMoves, bPath, bRuns, bArena, bArenaTop, bColTop, bReadOff, bBad;
	GaUmapSchedule sched;
	uint32_t debugFlags = 0;   // GA_DEBUG_FLAGS env: bit0 skip traceback (timing experiments only)
	double avgNodeLen = 32;    // mean node length of the uploaded graph (sizing heuristics)
	int forceS = 0;            // GA_STREAMS_PER_WARP env: override the streams-per-warp heuristic (tuning)
	int traceT = 0, traceP = 1; // GA_TRACE_T / GA_TRACE_P env: streams per warp / 32-column passes per window of the traceback kernel (tuning)
	int smCount = 148;
	int warpsPerSm = 20;       // resident warps of ga_align_kernel per SM (occupancy query)
	// pinned host staging (grow-only): parts for H2D, stream results + trace arena for D2H
	struct Pinned
	{
		void* ptr = nullptr;
		size_t cap = 0;
		void* ensure(size_t bytes)
		{
			if (bytes <= cap) return ptr;
			if (ptr) cudaFreeHost(ptr);
			ptr = nullptr;
			cap = 0;
			size_t want = bytes + bytes / 8 + 4096;
			GA_CUDA(cudaHostAlloc(&ptr, want, cudaHostAllocDefault));
			cap = want;
			return ptr;
		}
		void release()
		{
			if (ptr) cudaFreeHost(ptr);
			ptr = nullptr;
			cap = 0;
		}
	};
	Pinned pinParts, pinOuts, pinArena, pinSmall, pinReadOff;
	size_t budgetBytes = 0;    // device bytes a batch may use (FreeDeviceBytes), 0 = not queried yet
	cudaEvent_t evWait = nullptr;      // blocking-sync event of waitStream
	cudaEvent_t evKernel[4] = { nullptr, nullptr, nullptr, nullptr };   // around the three kernels of the last launch sequence
	bool evRecorded = false;
};

struct StagedBatch
{
	std::vector<ga_stream_in> sorted;      // streams in launch order (longest first)
	std::vector<uint32_t> perm;            // launch index -> caller's index
	ga_caps caps;
	ScratchPtrs sp;
	int b = 0, B = 0;
	uint64_t arenaCap = 0;
	size_t nWarps = 0;
	int launches = 0;
	// host copy of the inputs, kept for retries
	const uint8_t* hostParts = nullptr;
	size_t hostPartsBytes = 0;
	int capScale = 1;
	int S = 32;            // streams per warp
	bool smemScratch = false;   // small-band mode: per-slice scratch in shared memory
	bool noReplay = false;      // -B ramp: forward pass without the replay of the reference's stale sqrt checkpoints (second try of a stream, see FinishStaged)
	size_t peqWords = 0;
	uint64_t colPoolCap = 0;
	size_t nReads = 0;     // reads whose characters the run validates (SetReadRanges), 0 = none
};

static GaHmmTables makeHmmTables()
{
	// Same expressions, same association, same libm as the reference (AlignmentCorrectnessEstimation.cpp:6-26,51-59,
	// 81-85); the device only adds and compares these doubles.
	GaHmmTables t;
	const double correctMismatch = log(0.2);
	const double correctMatch = log(1.0 - 0.2);
	const double falseMismatch = log(0.5);
	const double falseMatch = log(1.0 - 0.5);
	std::vector<double> logFactorials;
	logFactorials.push_back(0);
	for (int i = 1; i <= 64; i++) logFactorials.push_back(logFactorials.back() + log(i));
	for (int m = 0; m <= 64; m++)
	{
		double chooseresult = logFactorials[64] - logFactorials[m] - logFactorials[64 - m];
		t.correctMul[m] = chooseresult + m * correctMismatch + (64 - m) * correctMatch;
		t.falseMul[m] = chooseresult + m * falseMismatch + (64 - m) * falseMatch;
	}
	t.f2c = log(0.00001);
	t.f2f = log(1.0 - 0.00001);
	t.c2f = log(0.000000000000001);
	t.c2c = log(1.0 - 0.000000000000001);
	t.startCorrect = log(0.8);
	t.startFalse = log(0.2);
	return t;
}

static GaUmapSchedule probeUmapSchedule(size_t maxElems)
{
	// bucket-count growth of the very std::unordered_map the reference uses for a slice's node map
	GaUmapSchedule sch;
	sch.n = 0;
	std::unordered_map<size_t, int> probe;
	size_t last = probe.bucket_count();
	for (size_t k = 1; k <= maxElems && sch.n < 48; k++)
	{
		probe[k] = 0;
		if (probe.bucket_count() != last)
		{
			last = probe.bucket_count();
			sch.threshold[sch.n] = (uint32_t)k;
			sch.buckets[sch.n] = (uint32_t)last;
			sch.n++;
		}
	}
	return sch;
}

DeviceCtx* CreateDevice(int device)
{
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) throw std::runtime_error(std::string("graphaligner_b200: no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU path");
	if (device < 0 || device >= count) throw std::runtime_error("graphaligner_b200: device index out of range");
	GA_CUDA(cudaSetDevice(device));
	DeviceCtx* ctx = new DeviceCtx();
	ctx->device = device;
	GA_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
	GaHmmTables hmm = makeHmmTables();
	GA_CUDA(cudaMemcpyToSymbol(c_hmm, &hmm, sizeof(hmm)));
	ctx->sched = probeUmapSchedule(70000);
	if (const char* f = getenv("GA_DEBUG_FLAGS")) ctx->debugFlags = (uint32_t)atoi(f);
	if (const char* f = getenv("GA_STREAMS_PER_WARP")) ctx->forceS = atoi(f);
	if (const char* f = getenv("GA_TRACE_T")) ctx->traceT = atoi(f);
	if (const char* f = getenv("GA_TRACE_P")) ctx->traceP = atoi(f);
	{
		cudaDeviceProp prop;
		GA_CUDA(cudaGetDeviceProperties(&prop, device));
		ctx->smCount = prop.multiProcessorCount;
		int blocks = 0;
#ifndef GA_HOSTSIM
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, ga_forward_kernel<4, false>, 64, 0) == cudaSuccess && blocks > 0) ctx->warpsPerSm = blocks * 2;
#else
		(void)blocks;
		ctx->forceS = 1;   // the emulation runs one stream per "warp"
#endif
	}
	GA_CUDA(cudaMemcpyToSymbol(c_sched, &ctx->sched, sizeof(GaUmapSchedule)));
	for (auto& e : ctx->evKernel) GA_CUDA(cudaEventCreate(&e));
#ifndef GA_HOSTSIM
	if (getenv("GA_SPIN_WAIT") == nullptr) GA_CUDA(cudaEventCreateWithFlags(&ctx->evWait, cudaEventBlockingSync | cudaEventDisableTiming));
#endif
	return ctx;
}

void DestroyDevice(DeviceCtx* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	Buffer* all[] = { &ctx->gNodeStart, &ctx->gSeq, &ctx->gInOff, &ctx->gInAdj, &ctx->gOutOff, &ctx->gOutAdj, &ctx->gNodeRec, &ctx->gChunks, &ctx->gNodeIdRev, &ctx->bParts, &ctx->bIn, &ctx->bOut, &ctx->bWd, &ctx->bTiny, &ctx->bHash,
		&ctx->bHeap, &ctx->bNodeTmp, &ctx->bUbkt, &ctx->bHdr, &ctx->bHn, &ctx->bColVV, &ctx->bColS, &ctx->bPeq, &ctx->bPeqAux, &ctx->bPeqOff, &ctx->bMoves, &ctx->bPath, &ctx->bRuns, &ctx->bArena, &ctx->bArenaTop, &ctx->bColTop, &ctx->bReadOff, &ctx->bBad };
	for (Buffer* b : all) b->release();
	ctx->pinParts.release();
	ctx->pinOuts.release();
	ctx->pinArena.release();
	ctx->pinSmall.release();
	ctx->pinReadOff.release();
	for (auto& e : ctx->evKernel) if (e) cudaEventDestroy(e);
	if (ctx->evWait) cudaEventDestroy(ctx->evWait);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

const std::string& LastError(DeviceCtx* ctx) { return ctx->lastError; }
void SetError(DeviceCtx* ctx, const std::string& msg) { ctx->lastError = msg; }
void* DeviceStream(DeviceCtx* ctx) { return (void*)ctx->stream; }
// Waits for the context's stream WITHOUT spinning: an event with cudaEventBlockingSync puts the thread to sleep.  The waits of
// a batch last milliseconds (the kernels, the D2H), a box runs a lane or two per GPU, and the cores are needed by the other
// lanes' host work - a spinning cudaStreamSynchronize per lane costs 16 of 32 cores on an 8-GPU box.
static void waitStream(DeviceCtx* ctx)
{
#ifdef GA_HOSTSIM
	GA_CUDA(cudaStreamSynchronize(ctx->stream));
#else
	if (!ctx->evWait) { GA_CUDA(cudaStreamSynchronize(ctx->stream)); return; }
	GA_CUDA(cudaEventRecord(ctx->evWait, ctx->stream));
	GA_CUDA(cudaEventSynchronize(ctx->evWait));
#endif
}

void SyncDevice(DeviceCtx* ctx)
{
	GA_CUDA(cudaSetDevice(ctx->device));
	waitStream(ctx);
}
size_t GraphBytesOnDevice(DeviceCtx* ctx) { return ctx->graphBytes; }

double MeasureInt32Peak(DeviceCtx* ctx)
{
	GA_CUDA(cudaSetDevice(ctx->device));
	cudaDeviceProp prop;
	GA_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
	ctx->bArenaTop.ensure(64);
	const int iters = 4096, threads = 256, blocks = prop.multiProcessorCount * 16;
	cudaEvent_t e0, e1;
	GA_CUDA(cudaEventCreate(&e0));
	GA_CUDA(cudaEventCreate(&e1));
	double best = 0;
	for (int rep = 0; rep < 5; rep++)
	{
		GA_CUDA(cudaEventRecord(e0, ctx->stream));
#ifndef GA_HOSTSIM
		ga_int32_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((uint32_t*)ctx->bArenaTop.ptr, iters);
#endif
		GA_CUDA(cudaEventRecord(e1, ctx->stream));
		GA_CUDA(cudaEventSynchronize(e1));
		float ms = 0;
		GA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		// 8 unrolled rounds x 8 chains; a LOP3 chain step is 1 instruction, an IADD3 chain step is 1 instruction
		double ops = (double)blocks * threads * iters * 8.0 * 8.0;
		double rate = ops / (ms * 1e-3);
		if (rep > 0 && rate > best) best = rate;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return best;
}

template <typename T>
static const T* uploadVec(DeviceCtx* ctx, Buffer& buf, const std::vector<T>& v)
{
	buf.ensure(std::max<size_t>(v.size() * sizeof(T), 16));
	GA_CUDA(cudaMemcpyAsync(buf.ptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
	ctx->graphBytes += v.size() * sizeof(T);
	return (const T*)buf.ptr;
}

void UploadGraph(DeviceCtx* ctx, const AlignmentGraph& graph)
{
	if (!graph.Finalized()) throw std::logic_error("UploadGraph: graph not finalized");
	GA_CUDA(cudaSetDevice(ctx->device));
	ctx->graphBytes = 0;
	ctx->view.nNodes = (uint32_t)graph.NodeSize();
	ctx->avgNodeLen = (double)graph.SizeInBp() / (double)std::max<size_t>(1, graph.NodeSize());
	ctx->view.nodeStart = uploadVec(ctx, ctx->gNodeStart, graph.NodeStarts());
	ctx->view.seq2 = uploadVec(ctx, ctx->gSeq, graph.Seq2());
	ctx->view.inOff = uploadVec(ctx, ctx->gInOff, graph.InOff());
	ctx->view.inAdj = uploadVec(ctx, ctx->gInAdj, graph.InAdj());
	ctx->view.outOff = uploadVec(ctx, ctx->gOutOff, graph.OutOff());
	ctx->view.outAdj = uploadVec(ctx, ctx->gOutAdj, graph.OutAdj());
	{
		// fixed-width node records and chunked sequence for the small-band forward kernel (ga_types.h)
		const size_t n = graph.NodeSize();
		const std::vector<uint64_t>& ns = graph.NodeStarts();
		const std::vector<uint32_t>& seq2 = graph.Seq2();
		std::vector<ga_node_rec> recs(n);
		uint64_t chunkTop = 0;
		for (size_t i = 0; i < n; i++) { recs[i].seqChunk = (uint32_t)chunkTop; chunkTop += (ns[i + 1] - ns[i] + 63) / 64; }
		if (chunkTop >= 0xffffffffull) throw std::runtime_error("graph too large for 32-bit sequence chunk indices");
		std::vector<uint32_t> chunks((size_t)chunkTop * 4 + 4, 0);
		ParallelFor(n, [&](size_t i) {
			ga_node_rec& r = recs[i];
			const uint64_t len = ns[i + 1] - ns[i];
			const uint32_t inDeg = graph.InOff()[i + 1] - graph.InOff()[i], outDeg = graph.OutOff()[i + 1] - graph.OutOff()[i];
			r.lenDeg = (uint32_t)std::min<uint64_t>(len, 0xffffffu) | (std::min(inDeg, 15u) << 24) | (std::min(outDeg, 15u) << 28);
			r.inOff = graph.InOff()[i];
			r.outOff = graph.OutOff()[i];
			for (uint32_t k = 0; k < 2; k++)
			{
				r.in[k] = k < inDeg ? graph.InAdj()[r.inOff + k] : 0xffffffffu;
				r.out[k] = k < outDeg ? graph.OutAdj()[r.outOff + k] : 0xffffffffu;
			}
			uint32_t* dst = chunks.data() + (size_t)r.seqChunk * 4;
			for (uint64_t k = 0; k < len; k++)
			{
				const uint64_t w = ns[i] + k;
				const uint32_t base = (seq2[w >> 4] >> ((uint32_t)(w & 15) * 2)) & 3u;
				dst[k >> 4] |= base << ((uint32_t)(k & 15) * 2);
			}
		});
		{
			std::vector<long long> idRev(n);
			for (size_t i = 0; i < n; i++) idRev[i] = ((long long)graph.NodeID(i) << 1) | (graph.Reverse(i) ? 1 : 0);
			ctx->view.nodeIdRev = uploadVec(ctx, ctx->gNodeIdRev, idRev);
			GA_CUDA(cudaStreamSynchronize(ctx->stream));
		}
		// one allocation for both, so that one L2 access-policy window covers them
		const size_t recBytes = (recs.size() * sizeof(ga_node_rec) + 255) / 256 * 256, chunkBytes = chunks.size() * sizeof(uint32_t);
		ctx->gNodeRec.ensure(recBytes + chunkBytes);
		GA_CUDA(cudaMemcpyAsync(ctx->gNodeRec.ptr, recs.data(), recs.size() * sizeof(ga_node_rec), cudaMemcpyHostToDevice, ctx->stream));
		GA_CUDA(cudaMemcpyAsync((uint8_t*)ctx->gNodeRec.ptr + recBytes, chunks.data(), chunkBytes, cudaMemcpyHostToDevice, ctx->stream));
		ctx->graphBytes += recBytes + chunkBytes;
		ctx->view.nodeRec = (const ga_node_rec*)ctx->gNodeRec.ptr;
		ctx->view.seqChunks = (const uint32_t*)((uint8_t*)ctx->gNodeRec.ptr + recBytes);
		GA_CUDA(cudaStreamSynchronize(ctx->stream));   // the vectors go out of scope
#ifndef GA_HOSTSIM
		// The node records and sequence chunks are what every band selection waits for (a dependent load per node that enters a
		// band): keep them in L2 across the gigabytes of history a launch streams through it.  Best effort - a device without
		// a persisting carve-out just runs without the window.
		{
			int maxPersist = 0, maxWindow = 0;
			cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
			cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
			const size_t want = recBytes + chunkBytes;
			if (maxPersist > 0 && maxWindow > 0 && getenv("GA_NO_L2_WINDOW") == nullptr)
			{
				const size_t persist = std::min<size_t>(want, (size_t)maxPersist / 2);   // half of what the device allows: the rest of L2 stays a normal cache
				const size_t window = std::min<size_t>(want, (size_t)maxWindow);
				if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess)
				{
					cudaStreamAttrValue attr;
					memset(&attr, 0, sizeof(attr));
					attr.accessPolicyWindow.base_ptr = ctx->gNodeRec.ptr;
					attr.accessPolicyWindow.num_bytes = window;
					attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)persist / (double)window);
					attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
					attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
					if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
				}
				else cudaGetLastError();
			}
		}
#endif
	}
	GA_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->hasGraph = true;
	ctx->budgetBytes = 0;
}

static uint32_t nextPow2(uint32_t v)
{
	uint32_t p = 1;
	while (p < v) p <<= 1;
	return p;
}

static ga_caps defaultCaps(int b, int B, int scale)
{
	int bw = std::max(b, B);
	ga_caps c;
	memset(&c, 0, sizeof(c));
	// a band holds the kept nodes (scores within bw of the minimum) plus everything within bw+64 bp downstream
	uint32_t cols = (uint32_t)(4 * (bw + 64) + 256);
	c.maxCols = std::max<uint32_t>(1024, cols * 2) * scale;
	c.maxNodes = std::min<uint32_t>(60000, std::max<uint32_t>(128, c.maxCols / 8));   // maxCols carries the scale already
	c.hashSize = nextPow2(c.maxNodes * 2);
	c.maxQueue = c.maxNodes * 8;
	return c;
}

// Whole-node banding: a band holds the kept nodes and their successors, so on a graph of long nodes (unchopped unitigs) a
// slice spans a few node lengths whatever the bandwidth - up to the reference's 200 000-bp switch to its alternate method.
// The capacities start at the scale that holds three average nodes instead of climbing there through overflow re-runs.
static int initialCapScale(const DeviceCtx* ctx, int b, int B)
{
	const double need = std::min<double>((double)GA_ALT_CUTOFF + 1024.0, 3.0 * ctx->avgNodeLen + 512.0);
	int scale = 1;
	while (scale < 1024 && (double)defaultCaps(b, B, scale).maxCols < need) scale *= 4;
	return scale;
}

StagedBatch* StageStreams(DeviceCtx* ctx, const std::vector<ga_stream_in>& streams, const uint8_t* parts, size_t partsBytes, int initialBandwidth, int rampBandwidth, BatchStats* stats,
	bool partsOnDevice = false)
{
	if (!ctx->hasGraph) throw std::logic_error("StageStreams: no graph uploaded to this device");
	GA_CUDA(cudaSetDevice(ctx->device));
	StagedBatch* sb = new StagedBatch();
	sb->b = initialBandwidth;
	sb->capScale = initialCapScale(ctx, initialBandwidth, rampBandwidth);
	sb->B = rampBandwidth;
	sb->hostParts = parts;
	sb->hostPartsBytes = partsBytes;
	// the big transfer first: sorting and the layout below run while it is in flight
	if (!partsOnDevice)
	{
		ctx->bParts.ensure(partsBytes + 64);
		if (partsBytes) GA_CUDA(cudaMemcpyAsync(ctx->bParts.ptr, parts, partsBytes, cudaMemcpyHostToDevice, ctx->stream));
	}
	const size_t n = streams.size();
	// longest streams first: the 32 streams of a warp iterate as long as the longest of them
	sb->perm.resize(n);
	std::iota(sb->perm.begin(), sb->perm.end(), 0u);
	std::stable_sort(sb->perm.begin(), sb->perm.end(), [&](uint32_t a, uint32_t b2) { return streams[a].partLen > streams[b2].partLen; });
	sb->sorted.resize(n);
	for (size_t i = 0; i < n; i++) sb->sorted[i] = streams[sb->perm[i]];
	return sb;
}

static int pickStreamsPerWarp(DeviceCtx* ctx, size_t nStreams)
{
	// enough warps to give every SM sub-partition several to switch between; full warps once the batch is large
	if (ctx->forceS > 0) return ctx->forceS;
	// as many warps as fit on the GPU in ONE wave (a second, nearly empty wave would double the run time of a kernel
	// whose warps all take about equally long), but never fewer than 4 streams per warp
	const size_t resident = (size_t)ctx->smCount * ctx->warpsPerSm;
	int S = 4;
	while (S < 32 && (nStreams + S - 1) / S > resident) S <<= 1;
	return S;
}

// Small-band kernel: a launch lasts as long as one warp's chain of slices, and a warp that shares its SM sub-partition
// with a second one is the slower for it.  So: the fewest streams per warp (16..32, not only powers of two - the lane
// interleave of every buffer is just a stride) that still give every warp a sub-partition of its own; a batch too large
// for that runs in waves of 16-stream warps, five warps of state per SM.
static int pickFastStreamsPerWarp(DeviceCtx* ctx, size_t nStreams)
{
	if (ctx->forceS > 0) return ctx->forceS;
	const size_t slots = (size_t)ctx->smCount * 4;
	// (a small batch runs with few streams per warp: the chain of a warp is the longest of its lanes' chains plus what their
	// divergent parts serialise, so fewer lanes per warp make a shorter launch as long as every warp still has a sub-partition)
	static const int candidates[] = { 4, 6, 8, 12, 16, 17, 18, 20, 24, 28, 32 };
	for (int S : candidates)
	{
		if ((nStreams + S - 1) / S <= slots) return S;
	}
	return 16;
}

static void layoutAndUpload(DeviceCtx* ctx, StagedBatch* sb, BatchStats* stats)
{
	const size_t n = sb->sorted.size();
	const int scale = sb->capScale;
	sb->caps = defaultCaps(sb->b, sb->B, scale);
	ga_caps& caps = sb->caps;
	{
		// small-band kernel (ga_fast.cuh): a graph whose bands hold a handful of nodes, fixed bandwidth; a stream that outgrows
		// the kernel's limits reports an overflow and is re-run by the general kernel (capScale > 1)
		const double bandNodes = 2.0 * (std::max(sb->b, sb->B) + 64) / std::max(1.0, ctx->avgNodeLen) + 2;
		// no -B ramp: a redo needs the general kernel's history rewind (slice 0 still runs with B, as in the reference)
		sb->smemScratch = scale == 1 && bandNodes <= 10 && sb->B <= sb->b && getenv("GA_NO_SMEM") == nullptr;
	}
	sb->S = sb->smemScratch ? pickFastStreamsPerWarp(ctx, n) : pickStreamsPerWarp(ctx, n);
	const size_t S = (size_t)sb->S;
	const size_t nWarps = (n + S - 1) / S;
	sb->nWarps = nWarps;
	const int bw = std::max(sb->b, sb->B);
	// expected band per slice: ~2 x (bandwidth + 64) columns plus a node or two of slack; nodes from the graph's mean node length
	const double avgNodeLen = std::max(1.0, ctx->avgNodeLen);
	const uint64_t colsGuess = (uint64_t)((2 * (bw + 64) + 2 * std::min(avgNodeLen, 256.0) + 32) * scale);
	const uint64_t nodesGuess = (uint64_t)(colsGuess / avgNodeLen * 1.5 + 8);
	std::vector<WarpDesc> wds(nWarps);
	std::vector<uint64_t> peqOff(n);
	uint64_t colTop = 0, hdrTop = 0, totalSlices = 0, hnTop = 0, movesTop = 0, pathTop = 0, runsTop = 0, peqTop = 0;
	sb->arenaCap = 0;
	for (size_t w = 0; w < nWarps; w++)
	{
		uint32_t maxLen = 0;
		for (size_t i = w * S; i < std::min(n, w * S + S); i++)
		{
			maxLen = std::max(maxLen, sb->sorted[i].partLen);
			// moves (2 bits each, ~1.2 per row) + crossed nodes + runs; nodes are assumed >= 4 bp on average, the retry path covers the rest
			sb->arenaCap += (uint64_t)sb->sorted[i].partLen * 3 / 16 + ((uint64_t)(sb->sorted[i].partLen * 1.3 / avgNodeLen * 2.0 * scale) + 66) * (1 + GA_MAP_WORDS) + 16;
			peqOff[i] = peqTop;
			peqTop += (uint64_t)(sb->sorted[i].partLen / 64) * 2;
		}
		uint32_t nslices = maxLen / 64;
		WarpDesc& d = wds[w];
		d.maxSlices = nslices;
		d.histNodes = (uint32_t)std::min<uint64_t>(0xfffffff0u, (uint64_t)nslices * nodesGuess + caps.maxNodes + 8);
		colTop += (uint64_t)nslices * colsGuess + 64;
		totalSlices += nslices;
		d.maxMoves = maxLen * 3 + 256;
		d.maxPathNodes = (uint32_t)std::min<uint64_t>(0x7fffffffu, (uint64_t)(maxLen * 1.3 / avgNodeLen * 2.0 * scale) + 64);
		d.maxRuns = d.maxPathNodes + 2;
		d.hdrBase = hdrTop;
		// (-B ramp: the checkpoint copies behind the slice headers, ga_core.cuh GA_CP_SLOTS)
		hdrTop += (uint64_t)(sb->B > sb->b ? GA_HDR_SLOTS_RAMP(d.maxSlices) : d.maxSlices) * GA_HDR_WORDS * S;
		d.hnBase = hnTop;
		hnTop += (uint64_t)d.histNodes * GA_HN_WORDS * S;
		d.movesBase = movesTop;
		movesTop += (uint64_t)(d.maxMoves / 16 + 1) * S;
		d.pathBase = pathTop;
		pathTop += (uint64_t)d.maxPathNodes * S;
		d.runsBase = runsTop;
		runsTop += (uint64_t)d.maxRuns * GA_RUN_WORDS * S;
	}
	sb->peqWords = peqTop;
	uint32_t ubktSize = 13;
	for (uint32_t i = 0; i < ctx->sched.n; i++)
	{
		ubktSize = ctx->sched.buckets[i];
		if (ctx->sched.buckets[i] >= caps.maxNodes) break;
	}
	if (sb->smemScratch)
	{
		// the small-band kernel's unordered_map emulation has 32 buckets for up to GAF_NODES keys
		for (uint32_t i = 0; i < ctx->sched.n; i++)
		{
			if (ctx->sched.threshold[i] <= GAF_NODES && ctx->sched.buckets[i] > 32) throw std::logic_error("unordered_map bucket schedule does not fit the small-band kernel's scratch");
		}
	}
	ubktSize = std::max(ubktSize, caps.maxNodes);
	ctx->bIn.ensure(n * sizeof(ga_stream_in));
	ctx->bOut.ensure(n * sizeof(ga_stream_out));
	ctx->bWd.ensure(nWarps * sizeof(WarpDesc));
	ctx->bTiny.ensure(nWarps * 3 * caps.maxCols * S * sizeof(uint32_t));
	ctx->bHash.ensure(nWarps * 2 * (size_t)caps.hashSize * S * sizeof(uint64_t));
	ctx->bHeap.ensure(nWarps * (size_t)caps.maxQueue * S * sizeof(uint64_t));
	ctx->bNodeTmp.ensure(nWarps * 10 * (size_t)caps.maxNodes * S * sizeof(uint32_t));
	ctx->bUbkt.ensure(nWarps * (size_t)ubktSize * S * sizeof(uint32_t));
	ctx->bHdr.ensure(hdrTop * sizeof(uint32_t));
	ctx->bHn.ensure(hnTop * sizeof(uint32_t));
	colTop += caps.maxCols;
	if (colTop >= 0xffffffffull) throw std::runtime_error("batch too large for one launch: column-history pool would exceed 2^32 columns; split the batch");
	sb->colPoolCap = colTop;
	ctx->bColVV.ensure(colTop * S * sizeof(uint4));
	ctx->bColS.ensure(colTop * S * sizeof(uint32_t));
	ctx->bColTop.ensure(sizeof(unsigned long long));
	ctx->bPeq.ensure(std::max<uint64_t>(peqTop, 1) * sizeof(uint4));
	ctx->bPeqAux.ensure(std::max<uint64_t>(peqTop / 2, 1) * sizeof(uint32_t));
	ctx->bPeqOff.ensure(n * sizeof(uint64_t));
	ctx->bMoves.ensure(movesTop * sizeof(uint32_t));
	ctx->bPath.ensure(pathTop * sizeof(uint32_t));
	ctx->bRuns.ensure(runsTop * sizeof(uint32_t));
	ctx->bArena.ensure(sb->arenaCap * sizeof(uint32_t));
	ctx->bArenaTop.ensure(sizeof(unsigned long long));
	sb->sp.tiny = (uint32_t*)ctx->bTiny.ptr;
	sb->sp.hash = (uint64_t*)ctx->bHash.ptr;
	sb->sp.heap = (uint64_t*)ctx->bHeap.ptr;
	sb->sp.nodeTmp = (uint32_t*)ctx->bNodeTmp.ptr;
	sb->sp.ubkt = (uint32_t*)ctx->bUbkt.ptr;
	sb->sp.ubktSize = ubktSize;
	sb->sp.hdr = (uint32_t*)ctx->bHdr.ptr;
	sb->sp.histNode = (uint32_t*)ctx->bHn.ptr;
	sb->sp.colVV = (uint4*)ctx->bColVV.ptr;
	sb->sp.colS = (uint32_t*)ctx->bColS.ptr;
	sb->sp.colPoolTop = (unsigned long long*)ctx->bColTop.ptr;
	sb->sp.colPoolCap = colTop;
	sb->sp.peq = (const uint4*)ctx->bPeq.ptr;
	sb->sp.peqAux = (uint32_t*)ctx->bPeqAux.ptr;
	sb->sp.peqOff = (const uint64_t*)ctx->bPeqOff.ptr;
	sb->sp.moves = (uint32_t*)ctx->bMoves.ptr;
	sb->sp.pathNodes = (uint32_t*)ctx->bPath.ptr;
	sb->sp.runs = (uint32_t*)ctx->bRuns.ptr;
	// the small tables go through pinned staging so that nothing here waits for the parts upload still in flight
	{
		const size_t bIn = n * sizeof(ga_stream_in), bWd = nWarps * sizeof(WarpDesc), bPo = n * sizeof(uint64_t);
		const size_t oWd = (bIn + 255) / 256 * 256, oPo = oWd + (bWd + 255) / 256 * 256;
		uint8_t* pin = (uint8_t*)ctx->pinSmall.ensure(oPo + bPo);
		memcpy(pin, sb->sorted.data(), bIn);
		memcpy(pin + oWd, wds.data(), bWd);
		memcpy(pin + oPo, peqOff.data(), bPo);
		GA_CUDA(cudaMemcpyAsync(ctx->bIn.ptr, pin, bIn, cudaMemcpyHostToDevice, ctx->stream));
		GA_CUDA(cudaMemcpyAsync(ctx->bWd.ptr, pin + oWd, bWd, cudaMemcpyHostToDevice, ctx->stream));
		GA_CUDA(cudaMemcpyAsync(ctx->bPeqOff.ptr, pin + oPo, bPo, cudaMemcpyHostToDevice, ctx->stream));
	}
	if (stats) stats->h2dBytes += sb->hostPartsBytes + n * sizeof(ga_stream_in) + nWarps * sizeof(WarpDesc) + n * sizeof(uint64_t);
}

#ifndef GA_HOSTSIM
// small-band kernel: one warp per block, the warp's state in dynamic shared memory
template <int S>
static void launchFast(DeviceCtx* ctx, StagedBatch* sb)
{
	const size_t n = sb->sorted.size();
	// per device: the attribute belongs to the function on the current device
	static std::atomic<uint64_t> attrDone(0);
	const uint64_t bit = 1ull << (ctx->device & 63);
	if (!(attrDone.load() & bit))
	{
		GA_CUDA(cudaFuncSetAttribute(ga_fast_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GaFastShared<S>)));
		attrDone.fetch_or(bit);
	}
	ga_fast_kernel<S><<<(unsigned)sb->nWarps, 32, sizeof(GaFastShared<S>), ctx->stream>>>(ctx->view, sb->caps, sb->sp, (const WarpDesc*)ctx->bWd.ptr, (const ga_stream_in*)ctx->bIn.ptr,
		(uint32_t)n, sb->b, sb->B, ctx->debugFlags | (sb->noReplay ? 2u : 0u), (ga_stream_out*)ctx->bOut.ptr);
}

template <int S>
static void launchAlign(DeviceCtx* ctx, StagedBatch* sb)
{
	const size_t n = sb->sorted.size();
	const int threads = 64;
	const unsigned blocks = (unsigned)((sb->nWarps * 32 + threads - 1) / threads);
	const size_t smemBytes = (size_t)(threads / 32) * 4 * S * sizeof(unsigned long long);
	// -B ramp: the instantiation that also replays the reference's sqrt checkpoints (ga_run_stream<.., RAMP>)
	if (sb->B > sb->b) ga_forward_kernel<S, false, true><<<blocks, threads, smemBytes, ctx->stream>>>(ctx->view, sb->caps, sb->sp, (const WarpDesc*)ctx->bWd.ptr, (const ga_stream_in*)ctx->bIn.ptr,
		(uint32_t)n, sb->b, sb->B, ctx->debugFlags | (sb->noReplay ? 2u : 0u), (ga_stream_out*)ctx->bOut.ptr);
	else ga_forward_kernel<S, false><<<blocks, threads, smemBytes, ctx->stream>>>(ctx->view, sb->caps, sb->sp, (const WarpDesc*)ctx->bWd.ptr, (const ga_stream_in*)ctx->bIn.ptr,
		(uint32_t)n, sb->b, sb->B, ctx->debugFlags | (sb->noReplay ? 2u : 0u), (ga_stream_out*)ctx->bOut.ptr);
}

template <int T, int P, bool ALT = false>
static void launchTraceTP(DeviceCtx* ctx, StagedBatch* sb)
{
	const size_t n = sb->sorted.size();
	static std::atomic<uint64_t> attrDone(0);   // per device, as in launchFast
	const uint64_t bit = 1ull << (ctx->device & 63);
	if (!(attrDone.load() & bit))
	{
		GA_CUDA(cudaFuncSetAttribute(ga_trace_kernel<T, P, ALT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GaTraceShared<T, P>)));
		attrDone.fetch_or(bit);
	}
	ga_trace_kernel<T, P, ALT><<<(unsigned)((n + T - 1) / T), 32, sizeof(GaTraceShared<T, P>), ctx->stream>>>(ctx->view, sb->sp, (const WarpDesc*)ctx->bWd.ptr, (const ga_stream_in*)ctx->bIn.ptr, (uint32_t)n,
		(uint32_t)sb->S, (ga_stream_out*)ctx->bOut.ptr, (uint32_t*)ctx->bArena.ptr, (unsigned long long*)ctx->bArenaTop.ptr, (unsigned long long)sb->arenaCap);
}

static void launchTrace(DeviceCtx* ctx, StagedBatch* sb)
{
	// streams per warp: few, so that a batch is many warps (the walk is latency-bound).  Measured on the B200 (168 registers:
	// 12 warps per SM): 10 000 streams 5.3 ms at T = 6 (one wave), 7.0 at 8, 8.9 at 5 (two waves); 28 000 streams in waves
	// 13.2 ms at T = 8, 14.8 at 6, 18.1 at 16, 28.3 at 32.
	const size_t n = sb->sorted.size();
	// -B ramp: the walk that tells a slice's checkpoint instance from its re-computed one (ga_trace.cuh, ALT)
	if (sb->B > sb->b) { launchTraceTP<8, 1, true>(ctx, sb); return; }
	int T = ctx->traceT;
	if (T == 0)
	{
		// one wave (12 warps per SM) with as few streams per warp as that allows, 8 once the batch runs in waves anyway
		const size_t slots = (size_t)ctx->smCount * 12;
		T = 8;
		for (int t : { 2, 3, 4, 5, 6 }) { if ((n + t - 1) / t <= slots) { T = t; break; } }
	}
	const int P = ctx->traceP;
	if (T == 2) launchTraceTP<2, 1>(ctx, sb);
	else if (T == 3) launchTraceTP<3, 1>(ctx, sb);
	else if (T == 6) launchTraceTP<6, 1>(ctx, sb);
	else if (T == 5) launchTraceTP<5, 1>(ctx, sb);
	else if (T == 7) launchTraceTP<7, 1>(ctx, sb);
	else if (T == 4 && P == 1) launchTraceTP<4, 1>(ctx, sb);
	else if (T == 10) launchTraceTP<10, 1>(ctx, sb);
	else if (T == 12) launchTraceTP<12, 1>(ctx, sb);
	else if (T == 8 && P == 1) launchTraceTP<8, 1>(ctx, sb);
	else if (T == 8) launchTraceTP<8, 2>(ctx, sb);
	else if (T == 16 && P == 1) launchTraceTP<16, 1>(ctx, sb);
	else if (T == 16) launchTraceTP<16, 2>(ctx, sb);
	else if (T == 4) launchTraceTP<4, 2>(ctx, sb);
	else if (P == 1) launchTraceTP<32, 1>(ctx, sb);
	else launchTraceTP<32, 2>(ctx, sb);
}
#endif

int RunStaged(DeviceCtx* ctx, StagedBatch* sb)
{
	GA_CUDA(cudaSetDevice(ctx->device));
	const size_t n = sb->sorted.size();
	if (n == 0) return 0;
	// the node -> slot tables rely on stamps; clear what an earlier batch left behind
	GA_CUDA(cudaMemsetAsync(ctx->bHash.ptr, 0, sb->nWarps * 2 * (size_t)sb->caps.hashSize * sb->S * sizeof(uint64_t), ctx->stream));
	GA_CUDA(cudaMemsetAsync(ctx->bArenaTop.ptr, 0, sizeof(unsigned long long), ctx->stream));
	GA_CUDA(cudaMemsetAsync(ctx->bColTop.ptr, 0, sizeof(unsigned long long), ctx->stream));
#ifdef GA_HOSTSIM
	hostsim_peq((const ga_stream_in*)ctx->bIn.ptr, (const uint64_t*)ctx->bPeqOff.ptr, (const uint8_t*)ctx->bParts.ptr, (uint32_t)n, (uint4*)ctx->bPeq.ptr, (uint32_t*)ctx->bPeqAux.ptr);
	for (size_t r = 0; r < sb->nReads; r++)
	{
		const uint64_t* ro = (const uint64_t*)ctx->bReadOff.ptr;
		uint32_t any = 0;
		for (uint64_t i = ro[r]; i < ro[r + 1]; i++) any |= ga_iupac_mask(((const uint8_t*)ctx->bParts.ptr)[i]) == 0 ? 1u : 0u;
		((uint32_t*)ctx->bBad.ptr)[r] = any;
	}
	hostsim_align(ctx->view, sb->caps, sb->sp, (const WarpDesc*)ctx->bWd.ptr, (const ga_stream_in*)ctx->bIn.ptr, (const uint8_t*)ctx->bParts.ptr, (uint32_t)n, sb->b, sb->B, ctx->debugFlags | (sb->noReplay ? 2u : 0u),
		sb->smemScratch, (ga_stream_out*)ctx->bOut.ptr, (uint32_t*)ctx->bArena.ptr, (unsigned long long*)ctx->bArenaTop.ptr, (unsigned long long)sb->arenaCap);
#else
	// device time of each kernel of the launch sequence: events on the stream (read in FinishStaged; GA_KERNEL_TIMING prints them)
	GA_CUDA(cudaEventRecord(ctx->evKernel[0], ctx->stream));
	{
		ga_peq_kernel<<<(unsigned)n, 128, 0, ctx->stream>>>((const ga_stream_in*)ctx->bIn.ptr, (const uint64_t*)ctx->bPeqOff.ptr, (const uint8_t*)ctx->bParts.ptr, (uint32_t)n, (uint4*)ctx->bPeq.ptr,
			(uint32_t*)ctx->bPeqAux.ptr);
		GA_CUDA(cudaGetLastError());
		if (sb->nReads) ga_validate_kernel<<<(unsigned)sb->nReads, 128, 0, ctx->stream>>>((const uint8_t*)ctx->bParts.ptr, (const uint64_t*)ctx->bReadOff.ptr, (uint32_t)sb->nReads, (uint32_t*)ctx->bBad.ptr);
		GA_CUDA(cudaGetLastError());
	}
	GA_CUDA(cudaEventRecord(ctx->evKernel[1], ctx->stream));
	if (sb->smemScratch)
	{
		switch (sb->S)
		{
			case 32: launchFast<32>(ctx, sb); break;
			case 28: launchFast<28>(ctx, sb); break;
			case 24: launchFast<24>(ctx, sb); break;
			case 20: launchFast<20>(ctx, sb); break;
			case 18: launchFast<18>(ctx, sb); break;
			case 17: launchFast<17>(ctx, sb); break;
			case 16: launchFast<16>(ctx, sb); break;
			case 12: launchFast<12>(ctx, sb); break;
			case 8: launchFast<8>(ctx, sb); break;
			case 6: launchFast<6>(ctx, sb); break;
			case 4: launchFast<4>(ctx, sb); break;
			default: throw std::logic_error("unsupported streams-per-warp (small-band kernel)");
		}
	}
	else switch (sb->S)
	{
		case 32: launchAlign<32>(ctx, sb); break;
		case 16: launchAlign<16>(ctx, sb); break;
		case 8: launchAlign<8>(ctx, sb); break;
		case 4: launchAlign<4>(ctx, sb); break;
		case 2: launchAlign<2>(ctx, sb); break;
		case 1: launchAlign<1>(ctx, sb); break;
		default: throw std::logic_error("unsupported streams-per-warp");
	}
	GA_CUDA(cudaGetLastError());
	GA_CUDA(cudaEventRecord(ctx->evKernel[2], ctx->stream));
	launchTrace(ctx, sb);
	GA_CUDA(cudaEventRecord(ctx->evKernel[3], ctx->stream));
	ctx->evRecorded = true;
	static const bool kernelTiming = getenv("GA_KERNEL_TIMING") != nullptr;
	if (kernelTiming)
	{
		GA_CUDA(cudaEventSynchronize(ctx->evKernel[3]));
		float a = 0, b = 0, c = 0;
		cudaEventElapsedTime(&a, ctx->evKernel[0], ctx->evKernel[1]);
		cudaEventElapsedTime(&b, ctx->evKernel[1], ctx->evKernel[2]);
		cudaEventElapsedTime(&c, ctx->evKernel[2], ctx->evKernel[3]);
		fprintf(stderr, "[ga kernels] streams %zu S %d: peq %.3f ms, forward %.3f ms, trace %.3f ms\n", n, sb->S, a, b, c);
	}
#endif
	GA_CUDA(cudaGetLastError());
	const int launched = sb->nReads ? 4 : 3;   // match masks, bad-character check, forward DP, traceback
	sb->launches += launched;
	return launched;
}

static bool isOverflow(int32_t status)
{
	return status == GA_ERR_NODE_OVERFLOW || status == GA_ERR_COL_OVERFLOW || status == GA_ERR_QUEUE_OVERFLOW || status == GA_ERR_HIST_OVERFLOW || status == GA_ERR_TRACE_OVERFLOW;
}

// ---- pinned host blocks for results -------------------------------------------------------------------------------
namespace
{
struct PinnedPool
{
	std::mutex m;
	std::vector<std::pair<void*, size_t>> blocks;   // oldest first
	size_t cached = 0;
	size_t limit;
	PinnedPool()
	{
		limit = (size_t)1024 << 20;
		if (const char* e = getenv("GA_PINNED_CACHE_MB")) limit = (size_t)std::max(0ll, atoll(e)) << 20;
	}
};
PinnedPool& pinnedPool()
{
	static PinnedPool* p = new PinnedPool();   // never destroyed: results may be freed during static destruction
	return *p;
}

void* pinnedAcquire(size_t bytes, size_t& capOut)
{
	PinnedPool& pool = pinnedPool();
	{
		std::lock_guard<std::mutex> lock(pool.m);
		size_t best = pool.blocks.size();
		for (size_t i = 0; i < pool.blocks.size(); i++)
		{
			const size_t cap = pool.blocks[i].second;
			if (cap >= bytes && cap <= 2 * bytes + (8u << 20) && (best == pool.blocks.size() || cap < pool.blocks[best].second)) best = i;
		}
		if (best != pool.blocks.size())
		{
			void* p = pool.blocks[best].first;
			capOut = pool.blocks[best].second;
			pool.cached -= capOut;
			pool.blocks.erase(pool.blocks.begin() + best);
			return p;
		}
	}
	const size_t round = (size_t)4 << 20;
	capOut = (bytes + bytes / 8 + round - 1) / round * round;
	void* p = nullptr;
	GA_CUDA(cudaHostAlloc(&p, capOut, cudaHostAllocPortable));
	return p;
}

void pinnedRelease(void* p, size_t cap)
{
	if (!p) return;
	PinnedPool& pool = pinnedPool();
	std::vector<void*> drop;
	{
		std::lock_guard<std::mutex> lock(pool.m);
		pool.blocks.emplace_back(p, cap);
		pool.cached += cap;
		while (pool.cached > pool.limit && !pool.blocks.empty())
		{
			drop.push_back(pool.blocks.front().first);
			pool.cached -= pool.blocks.front().second;
			pool.blocks.erase(pool.blocks.begin());
		}
	}
	for (void* d : drop) cudaFreeHost(d);
}
}

void FinishStaged(DeviceCtx* ctx, StagedBatch* sb, RawBuffer<ga_stream_out>& outs, RawBuffer<uint32_t>& arena, BatchStats* stats, std::vector<uint8_t>* badChar, size_t* mapTail)
{
	if (mapTail) *mapTail = 0;
	GA_CUDA(cudaSetDevice(ctx->device));
	const size_t n = sb->sorted.size();
	outs.resize(n);
	arena.clear();
	if (n == 0) return;
	uint32_t* pinBad = nullptr;
	if (badChar && sb->nReads)
	{
		pinBad = (uint32_t*)ctx->pinReadOff.ensure((sb->nReads + 1) * sizeof(uint64_t));   // the offsets were uploaded long ago: the staging block is free again
		GA_CUDA(cudaMemcpyAsync(pinBad, ctx->bBad.ptr, sb->nReads * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	}
	// D2H through pinned staging, then a parallel copy into the caller's buffers
	ga_stream_out* pinOuts = (ga_stream_out*)ctx->pinOuts.ensure(n * sizeof(ga_stream_out) + sizeof(unsigned long long));
	unsigned long long* pinTop = (unsigned long long*)(pinOuts + n);
	const bool timing = getenv("GA_TIMING") != nullptr;
	auto tLast = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!timing) return;
		auto now = std::chrono::steady_clock::now();
		fprintf(stderr, "[ga timing]   d2h: %-22s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tLast).count());
		tLast = now;
	};
	GA_CUDA(cudaMemcpyAsync(pinOuts, ctx->bOut.ptr, n * sizeof(ga_stream_out), cudaMemcpyDeviceToHost, ctx->stream));
	GA_CUDA(cudaMemcpyAsync(pinTop, ctx->bArenaTop.ptr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
	waitStream(ctx);
	lap("stream records");
	if (stats && ctx->evRecorded)
	{
		float a = 0, b = 0, c = 0;
		if (cudaEventElapsedTime(&a, ctx->evKernel[0], ctx->evKernel[1]) == cudaSuccess && cudaEventElapsedTime(&b, ctx->evKernel[1], ctx->evKernel[2]) == cudaSuccess
			&& cudaEventElapsedTime(&c, ctx->evKernel[2], ctx->evKernel[3]) == cudaSuccess)
		{
			stats->peqMs += a; stats->forwardMs += b; stats->traceMs += c; stats->kernelMs += a + b + c;
		}
		ctx->evRecorded = false;
	}
	unsigned long long top = *pinTop;
	if (top > sb->arenaCap) top = sb->arenaCap;
	if (pinBad)
	{
		badChar->assign(sb->nReads, 0);
		for (size_t r = 0; r < sb->nReads; r++) (*badChar)[r] = pinBad[r] ? 1 : 0;
	}
	// the trace arena lands in a pinned block that the results then own (recycled through a process-wide pool): no
	// second pass over ~100 MB, no fresh pages to fault in
	size_t tailWords = 0;
	if (mapTail)
	{
		for (size_t i = 0; i < n; i++) if (pinOuts[i].nMapped == 0 && !isOverflow(pinOuts[i].status)) tailWords += (size_t)pinOuts[i].nRuns * GA_MAP_WORDS;
	}
	if (top || tailWords)
	{
		const size_t topAligned = ((size_t)top + 7) & ~(size_t)7;
		const size_t total = mapTail ? topAligned + tailWords : (size_t)top;
		size_t cap = 0;
		uint32_t* pin = (uint32_t*)pinnedAcquire(total * sizeof(uint32_t) + 16, cap);
		arena.adopt(pin, total, cap, pinnedRelease);
		if (mapTail) *mapTail = topAligned;
		lap("pinned block");
		if (top) GA_CUDA(cudaMemcpyAsync(pin, ctx->bArena.ptr, top * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	}
	for (size_t i = 0; i < n; i++) outs.data()[sb->perm[i]] = pinOuts[i];
	waitStream(ctx);
	lap("trace arena");
#ifdef GA_PHASE_TIMING
	{
		static const char* names[16] = { "slice start: lock-step wait", "in-degrees", "node start: store + rest", "word-step loop", "successors", "slice end: minima, HMM, headers, wait", "traceback", "trace start",
			"band: map order", "band: kept nodes", "band: heap walk", "node start: in-edges", "node start: candidates", "slice: masks, setup", "slice: pool alloc", "-" };
		double sum[16] = { 0 }, total = 0;
		for (size_t i = 0; i < n; i++) for (int k = 0; k < 16; k++) sum[k] += (double)pinOuts[i].phase[k];
		for (int k = 0; k < 16; k++) total += sum[k];
		fprintf(stderr, "[ga phases] mean cycles per stream: total %.0f\n", total / n);
		for (int k = 0; k < 15; k++) fprintf(stderr, "[ga phases]   %-40s %12.0f  %5.1f %%\n", names[k], sum[k] / n, 100.0 * sum[k] / total);
	}
#endif
	if (stats)
	{
		stats->d2hBytes += n * sizeof(ga_stream_out) + top * sizeof(uint32_t) + sizeof(top);
		stats->launches += sb->launches;
		sb->launches = 0;
	}
	// streams that ran out of scratch are re-run with larger capacities (never on the CPU); then (-B ramp) streams whose walk through
	// a stretch re-computed from a stale checkpoint failed - a band without the node the walk stands on, where the reference
	// reads a slice that is not there and crashes - are re-run without the replay: they report the alignment of their own
	// forward pass instead of a failure
	std::vector<uint32_t> again;
	for (int pass = 0; pass < 2; pass++)
	{
	again.clear();
	for (size_t i = 0; i < n; i++)
	{
		const ga_stream_out& o = outs.data()[i];
		const bool staleFailure = (o.rampRedos & GA_RAMP_STALE_BIT) != 0 && o.status != GA_OK && o.status != GA_EMPTY && !isOverflow(o.status);
		if (pass == 0 ? isOverflow(o.status) : staleFailure) again.push_back((uint32_t)i);
	}
	if (!again.empty() && (pass == 0 ? sb->capScale < 1024 : !sb->noReplay))   // 1024: capacities beyond the 200 000-column switch to the alternate method
	{
		std::vector<uint32_t> inv(n);
		for (size_t i = 0; i < n; i++) inv[sb->perm[i]] = (uint32_t)i;
		std::vector<ga_stream_in> sub(again.size());
		for (size_t k = 0; k < again.size(); k++) sub[k] = sb->sorted[inv[again[k]]];
		// NOTE: the retry reuses the context's device buffers, so the staged batch cannot be run again afterwards
		StagedBatch* retry = StageStreams(ctx, sub, sb->hostParts, sb->hostPartsBytes, sb->b, sb->B, stats);
		retry->capScale = pass == 0 ? sb->capScale * 4 : sb->capScale;
		retry->noReplay = sb->noReplay || pass == 1;
		RawBuffer<ga_stream_out> subOuts;
		RawBuffer<uint32_t> subArena;
		size_t subTail = 0;
		try
		{
			layoutAndUpload(ctx, retry, stats);
			RunStaged(ctx, retry);
			FinishStaged(ctx, retry, subOuts, subArena, stats, nullptr, mapTail ? &subTail : nullptr);
		}
		catch (...)
		{
			delete retry;
			throw;
		}
		delete retry;
		if (stats) stats->retries += again.size();
		RawBuffer<uint32_t> merged;
		uint64_t base = arena.size();
		if (mapTail)
		{
			// [records of the first launch | records of the re-run | room for both launches' host-written mappings]
			const size_t recA = *mapTail, recB = subTail, roomA = arena.size() - recA, roomB = subArena.size() - recB;
			merged.resize(recA + recB + roomA + roomB);
			if (recA) memcpy(merged.data(), arena.data(), recA * sizeof(uint32_t));
			if (recB) memcpy(merged.data() + recA, subArena.data(), recB * sizeof(uint32_t));
			base = recA;
			*mapTail = recA + recB;
		}
		else
		{
			merged.resize(arena.size() + subArena.size());
			if (arena.size()) memcpy(merged.data(), arena.data(), arena.size() * sizeof(uint32_t));
			if (subArena.size()) memcpy(merged.data() + arena.size(), subArena.data(), subArena.size() * sizeof(uint32_t));
		}
		arena.swap(merged);
		for (size_t k = 0; k < again.size(); k++)
		{
			outs.data()[again[k]] = subOuts.data()[k];
			outs.data()[again[k]].traceOff += base;
		}
	}
	}
}

size_t EstimateStreamBytes(DeviceCtx* ctx, size_t partLen, int bandwidth)
{
	// what layoutAndUpload allocates per stream, from the same terms: the per-slice scratch of the general layout (sized by
	// the capacities of the launch, which grow with the graph's node length), the column history, node lists, headers,
	// match words and the trace buffers
	const int scale = initialCapScale(ctx, bandwidth, 0);
	const ga_caps caps = defaultCaps(bandwidth, 0, scale);
	const double avgNodeLen = std::max(1.0, ctx->avgNodeLen);
	const double colsGuess = (2.0 * (bandwidth + 64) + 2.0 * std::min(avgNodeLen, 256.0) + 32) * scale;
	const double nodesGuess = colsGuess / avgNodeLen * 1.5 + 8;
	const double slices = (double)((partLen + 63) / 64);
	const double scratch = 3.0 * caps.maxCols * sizeof(uint32_t) + 2.0 * caps.hashSize * sizeof(uint64_t) + (double)caps.maxQueue * sizeof(uint64_t) + 10.0 * caps.maxNodes * sizeof(uint32_t)
		+ 2.0 * caps.maxNodes * sizeof(uint32_t);
	const double history = slices * (colsGuess * (sizeof(uint4) + sizeof(uint32_t)) + nodesGuess * GA_HN_WORDS * sizeof(uint32_t) + GA_HDR_WORDS * sizeof(uint32_t) + 36.0);
	const double trace = partLen * 1.5 + (partLen * 1.3 / avgNodeLen * 2.0 * scale + 66) * (1 + GA_MAP_WORDS + GA_RUN_WORDS) * sizeof(uint32_t) + 1024;
	return (size_t)((scratch + history + trace) * 1.15);
}

size_t FreeDeviceBytes(DeviceCtx* ctx)
{
	GA_CUDA(cudaSetDevice(ctx->device));
	// cudaMemGetInfo costs 1 - 40 ms per call on a busy box.  Free memory plus the context's own grow-only pools is what a
	// batch may use, and that sum does not change when the pools grow: it is queried once per uploaded graph.
	size_t freeB = ctx->budgetBytes;
	if (freeB == 0)
	{
	size_t totalB = 0;
	GA_CUDA(cudaMemGetInfo(&freeB, &totalB));
	Buffer* all[] = { &ctx->bParts, &ctx->bIn, &ctx->bOut, &ctx->bWd, &ctx->bTiny, &ctx->bHash, &ctx->bHeap, &ctx->bNodeTmp, &ctx->bUbkt, &ctx->bHdr, &ctx->bHn, &ctx->bColVV, &ctx->bColS, &ctx->bPeq, &ctx->bPeqAux,
		&ctx->bPeqOff, &ctx->bMoves, &ctx->bPath, &ctx->bRuns, &ctx->bArena };
	for (Buffer* b : all) freeB += b->cap;
	ctx->budgetBytes = freeB;
	}
	if (const char* e = getenv("GA_MEM_BUDGET_MB")) freeB = std::min<size_t>(freeB, (size_t)atoll(e) << 20);   // testing: force batch splitting
	return freeB;
}

void SetReadRanges(DeviceCtx* ctx, StagedBatch* sb, const std::vector<uint64_t>& readOff)
{
	if (readOff.size() < 2) return;
	GA_CUDA(cudaSetDevice(ctx->device));
	const size_t bytes = readOff.size() * sizeof(uint64_t);
	ctx->bReadOff.ensure(bytes);
	ctx->bBad.ensure((readOff.size() - 1) * sizeof(uint32_t));
	void* pin = ctx->pinReadOff.ensure(bytes);
	memcpy(pin, readOff.data(), bytes);
	GA_CUDA(cudaMemcpyAsync(ctx->bReadOff.ptr, pin, bytes, cudaMemcpyHostToDevice, ctx->stream));
	sb->nReads = readOff.size() - 1;
}

uint8_t* AllocPinnedParts(DeviceCtx* ctx, size_t bytes)
{
	GA_CUDA(cudaSetDevice(ctx->device));
	ctx->bParts.ensure(bytes + 64);   // its device twin, so that ranges can be uploaded while the rest is built
	return (uint8_t*)ctx->pinParts.ensure(bytes);
}

void EnsureDeviceParts(DeviceCtx* ctx, size_t bytes)
{
	GA_CUDA(cudaSetDevice(ctx->device));
	ctx->bParts.ensure(bytes + 64);
}

bool IsPinnedHost(const void* p)
{
#ifdef GA_HOSTSIM
	(void)p;
	return false;
#else
	cudaPointerAttributes attr;
	memset(&attr, 0, sizeof(attr));
	if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return attr.type == cudaMemoryTypeHost;
#endif
}

void UploadPartsRange(DeviceCtx* ctx, const uint8_t* parts, size_t offset, size_t bytes)
{
	if (bytes == 0) return;
	GA_CUDA(cudaSetDevice(ctx->device));
	GA_CUDA(cudaMemcpyAsync((uint8_t*)ctx->bParts.ptr + offset, parts + offset, bytes, cudaMemcpyHostToDevice, ctx->stream));
}

void FreeStaged(DeviceCtx* ctx, StagedBatch* sb)
{
	(void)ctx;
	delete sb;
}

// exposed for ga_device users: layout + upload happen at stage time
StagedBatch* StageAndUpload(DeviceCtx* ctx, const std::vector<ga_stream_in>& streams, const uint8_t* parts, size_t partsBytes, int b, int B, BatchStats* stats, bool partsOnDevice)
{
	StagedBatch* sb = StageStreams(ctx, streams, parts, partsBytes, b, B, stats, partsOnDevice);
	try
	{
		layoutAndUpload(ctx, sb, stats);
	}
	catch (...)
	{
		delete sb;
		throw;
	}
	return sb;
}

void ExecuteStreams(DeviceCtx* ctx, const std::vector<ga_stream_in>& streams, const uint8_t* parts, size_t partsBytes, const std::vector<uint64_t>& readOff, int initialBandwidth, int rampBandwidth,
	RawBuffer<ga_stream_out>& outs, RawBuffer<uint32_t>& arena, std::vector<uint8_t>& badChar, BatchStats* stats)
{
	StagedBatch* sb = StageAndUpload(ctx, streams, parts, partsBytes, initialBandwidth, rampBandwidth, stats);
	try
	{
		SetReadRanges(ctx, sb, readOff);
		RunStaged(ctx, sb);
		FinishStaged(ctx, sb, outs, arena, stats, &badChar);
	}
	catch (...)
	{
		delete sb;
		throw;
	}
	delete sb;
}

}
