// Plain-old-data types shared by the host C++ layer and the CUDA kernels.
// No torch / STL types here: these structs cross the extern "C" boundary (include/graphaligner_b200.h).
#ifndef GA_TYPES_H
#define GA_TYPES_H
#include <stdint.h>

// Device-resident view of the AlignmentGraph (reference AlignmentGraph.h:46-56), flattened:
//   nodeStart   prefix offsets in bp over the concatenated digraph sequence, nNodes+1 entries
//               (index 0 and nNodes-1 are the reference's 1-bp dummy nodes, AlignmentGraph.cpp:22-30,108-118)
//   seq2        2-bit packed bases (A=0,C=1,G=2,T=3), 16 per 32-bit word, indexed by global bp
//   inOff/inAdj, outOff/outAdj  CSR of in-/out-neighbours IN THE REFERENCE'S INSERTION ORDER
//               (AlignmentGraph.cpp:104-105) - the traceback tie-break depends on it
typedef struct ga_graph_view
{
	uint32_t nNodes;
	const uint64_t* nodeStart;
	const uint32_t* seq2;
	const uint32_t* inOff;
	const uint32_t* inAdj;
	const uint32_t* outOff;
	const uint32_t* outAdj;
	/* The same graph as fixed-width node records for the small-band forward kernel (ga_fast.cuh): one 32-byte record per
	 * node (two 128-bit loads of one sector) and the node sequences as 16-byte chunks of 64 two-bit bases, every node starting
	 * on a chunk (one 128-bit load per 64 columns).  Built by the device layer at upload; NULL on the host side. */
	const struct ga_node_rec* nodeRec;
	const uint32_t* seqChunks;    /* 4 words per chunk */
	const long long* nodeIdRev;   /* per node: digraph id << 1 | reverse flag (AlignmentGraph::NodeID / Reverse), for the mapping records */
} ga_graph_view;

/* len and degrees: len | min(inDeg, 15) << 24 | min(outDeg, 15) << 28.  in[] / out[]: the first two neighbours in the
 * reference's insertion order (0xffffffff = none); longer lists continue in inAdj / outAdj at inOff / outOff. */
typedef struct ga_node_rec
{
	uint32_t seqChunk;
	uint32_t lenDeg;
	uint32_t inOff;
	uint32_t outOff;
	uint32_t in[2];
	uint32_t out[2];
} ga_node_rec;
#define GA_REC_LEN(r) ((r) & 0xffffffu)
#define GA_REC_INDEG(r) (((r) >> 24) & 15u)
#define GA_REC_OUTDEG(r) ((r) >> 28)

// One DP stream = one direction of one (read, seed) pair: the padded read part aligned forward from a
// start node (reference getSplitAlignment, GraphAligner.h:2969-3024).
typedef struct ga_stream_in
{
	uint64_t seqOff;     // byte offset, in the batch's raw read bytes, of the part's first character (a backward part walks down from it)
	uint32_t partLen;    // padded length (multiple of 64) = reference sequence.size()
	uint32_t startNode;  // graph node index whose columns are all 0 in the initial slice
	uint32_t trimRows;   // trace positions with row >= trimRows are dropped (padding / DBG overlap, GraphAligner.h:3063-3066,3086-3089)
	uint32_t srcInfo;    // real (unpadded) length of the part | GA_SRC_SOLO | backward << 31 (backward: reverse complement of the read's prefix)
} ga_stream_in;
#define GA_SRC_LEN(x) ((x) & 0x3fffffffu)
#define GA_SRC_BACKWARD 0x80000000u
/* the stream is everything its read has (one seed at read position 0: forward part only, rows not shifted): the traceback
 * kernel writes the read's final 32-byte mapping records instead of the run records (ga_stream_out::nMapped) */
#define GA_SRC_SOLO 0x40000000u

enum
{
	GA_OK = 0,
	GA_EMPTY = 1,                // no slice survived removeWronglyAlignedEnd -> this direction failed
	GA_ERR_NODE_OVERFLOW = 2,    // band had more nodes than maxNodes (host retries with larger caps)
	GA_ERR_COL_OVERFLOW = 3,     // band had more columns than maxCols / column history exhausted
	GA_ERR_QUEUE_OVERFLOW = 4,
	GA_ERR_HIST_OVERFLOW = 5,    // node-list history exhausted
	GA_ERR_ALT_METHOD = 6,       // band >= 200000 bp: reference switches to calculateSliceAlternate (not built)
	GA_ERR_TRACE = 7,            // traceback found no predecessor (reference: assert(false); abort())
	GA_ERR_TRACE_OVERFLOW = 8,
	GA_ERR_INTERNAL = 9,
	GA_ERR_CYCLE_ITER = 10       // cyclic component did not converge within the iteration cap
};

#define GA_MAX_TIES 8

typedef struct ga_stream_out
{
	int32_t status;
	int32_t nSlices;        // slices retained after removeWronglyAlignedEnd (= bandwidthPerSlice.size())
	int32_t score;          // minScore of the last retained slice
	int32_t nSlicesRun;     // forward slices evaluated (incl. the one that triggered the early stop)
	uint64_t wordColumns;   // forward-pass word updates (SURVEY 8d "W")
	uint32_t endNode;       // trace start cell: node index, offset in node (row = 64*nSlices-1)
	uint32_t endOff;
	uint32_t nMoves;        // number of 2-bit moves in the trace (backward order)
	uint32_t nPathNodes;    // node indices crossed into, backward order
	uint32_t nRuns;         // maximal same-node runs of the trimmed trace, backward order, GA_RUN_WORDS words each
	uint32_t nPositions;    // trace positions left after trimming
	uint64_t traceOff;      // offset (in 32-bit words) of this stream's record in the trace arena
	uint32_t nTies;         // cells of the last retained slice tied at the minimum (incl. the chosen one)
	uint32_t cyclicSlices;  // slices whose band held a cyclic component
	uint32_t rampRedos;     // -B ramp: how often the stream went back and redid a stretch with the wide band (GraphAligner.h:2648-2719); GA_RAMP_STALE_BIT: a checkpoint of the abandoned pass was used
	uint32_t nMapped;       // > 0: the stream's record holds nMapped (= nRuns) GaDeviceMapping records where the runs would be, 32-byte aligned in the arena
	uint32_t tieNode[GA_MAX_TIES];
	uint32_t tieOff[GA_MAX_TIES];
#ifdef GA_PHASE_TIMING
	unsigned long long phase[16];  /* profiling builds: cycles per phase, see ga_core.cuh */
#endif
} ga_stream_out;

#define GA_RAMP_STALE_BIT 0x80000000u
#define GA_RUN_WORDS 5   /* node, firstOff, lastOff, firstRow, lastRow (first = smallest row) */

/* one vg::Mapping with its single Edit as the C ABI returns it (ga_mapping in include/graphaligner_b200.h, same layout) */
typedef struct GaDeviceMapping
{
	long long node_id;
	uint32_t offset;
	uint32_t rank;
	int32_t from_length;
	int32_t to_length;
	uint32_t read_start;
	uint32_t is_reverse;
} GaDeviceMapping;
#define GA_MAP_WORDS 8
/* words of padding in front of a stream's mapping records so that they start on a 32-byte boundary of the arena */
#define GA_MAP_PAD(wordOff) ((8u - (uint32_t)((wordOff) & 7u)) & 7u)

// moves (2 bits each, backward from the end cell)
enum { GA_MOVE_H = 0, GA_MOVE_D = 1, GA_MOVE_V = 2, GA_MOVE_END = 3 };

typedef struct ga_caps
{
	uint32_t maxNodes;      // band nodes per slice (per stream)
	uint32_t maxCols;       // band columns per slice (per stream)
	uint32_t hashSize;      // power of two, >= 2*maxNodes
	uint32_t maxQueue;      // heap / ready-queue entries
	uint32_t maxSlices;     // slice headers per stream
	uint32_t histNodes;     // node-list history entries per stream
	uint64_t warpCols;      // capacity of the shared column-history pool, in columns (x lanes x 32 B); must stay below 2^32
	uint32_t maxMoves;      // per-stream temporary trace capacity (moves)
	uint32_t maxPathNodes;
	uint32_t maxRuns;
} ga_caps;

#endif
