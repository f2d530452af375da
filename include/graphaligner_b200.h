/* graphaligner_b200 - C ABI of the B200-native GraphAligner hot path.
 *
 * Drop-in boundary for the reference's seeded alignment call
 *     AlignmentResult AlignOneWay(const AlignmentGraph&, const std::string& seq_id, const std::string& sequence,
 *                                 int initialBandwidth, int rampBandwidth, size_t dynamicRowStart,
 *                                 const std::vector<std::tuple<int,size_t,bool>>& seedHits)
 * (reference GraphAlignerWrapper.h:54, called from Aligner.cpp:140), batched: one call aligns many reads on one
 * GPU.  Plain pointers and sizes only; no C++ or torch types cross this boundary.  Return codes instead of
 * exceptions; a context is bound to one GPU and is not thread-safe, different contexts are independent.
 * There is no CPU path: ga_create fails when no CUDA device is usable.
 */
#ifndef GRAPHALIGNER_B200_H
#define GRAPHALIGNER_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ga_ctx ga_ctx;
typedef struct ga_graph ga_graph;
typedef struct ga_results ga_results;
typedef struct ga_staged ga_staged;

/* ---- context (one per GPU) ------------------------------------------------------------------------------- */
ga_ctx* ga_create(int device);                 /* NULL on failure; message via ga_global_error() */
void ga_destroy(ga_ctx* ctx);
const char* ga_last_error(const ga_ctx* ctx);  /* last error of this context ("" if none) */
const char* ga_global_error(void);             /* last error of a call that had no context */

/* ---- graph: replaces class AlignmentGraph (reference AlignmentGraph.h:25-43) ----------------------------- */
ga_graph* ga_graph_new(void);
/* AlignmentGraph::AddNode(nodeId, sequence, reverseNode), AlignmentGraph.cpp:44-91; ids are digraph ids (2*id, 2*id+1) */
int ga_graph_add_node(ga_graph* g, int32_t digraph_node_id, const char* sequence, size_t length, int reverse_node);
/* AlignmentGraph::AddEdgeNodeId, AlignmentGraph.cpp:93-106 */
int ga_graph_add_edge(ga_graph* g, int32_t digraph_from, int32_t digraph_to);
int ga_graph_set_dbg_overlap(ga_graph* g, int32_t overlap);   /* AlignmentGraph::DBGOverlap */
int ga_graph_finalize(ga_graph* g);                            /* AlignmentGraph::Finalize(64) */
/* DirectedGraph::StreamVGGraphFromFile / StreamGFAGraphFromFile (BigraphToDigraph.cpp:106-189) from memory:
 * n_nodes bidirected nodes (ids, sequences concatenated with n_nodes+1 offsets), n_edges edges.
 * gfa_overlap < 0: vg semantics; >= 0: GFA semantics with that edge overlap. */
ga_graph* ga_graph_from_bigraph(size_t n_nodes, const int64_t* ids, const char* sequences, const uint64_t* seq_offsets,
                                size_t n_edges, const int64_t* from, const uint8_t* from_start, const int64_t* to, const uint8_t* to_end,
                                int32_t gfa_overlap);
ga_graph* ga_graph_load_vg(const char* path);
ga_graph* ga_graph_load_gfa(const char* path);
void ga_graph_free(ga_graph* g);
size_t ga_graph_node_count(const ga_graph* g);   /* digraph nodes incl. the two dummy nodes */
size_t ga_graph_size_bp(const ga_graph* g);
size_t ga_graph_edge_count(const ga_graph* g);
/* copy the flattened graph (CSR + 2-bit sequence) into this context's GPU */
int ga_graph_upload(ga_ctx* ctx, const ga_graph* g);

/* ---- batch alignment -------------------------------------------------------------------------------------- */
typedef struct ga_batch
{
	size_t n_reads;
	const char* sequences;          /* all reads concatenated */
	const uint64_t* seq_offsets;    /* n_reads + 1 */
	const char* names;              /* all names concatenated (may be NULL) */
	const uint64_t* name_offsets;   /* n_reads + 1 (may be NULL) */
	const uint64_t* seed_offsets;   /* n_reads + 1: seeds of read i are [seed_offsets[i], seed_offsets[i+1]) */
	const int32_t* seed_node;       /* bigraph node id   (std::get<0> of the reference seed tuple) */
	const uint64_t* seed_pos;       /* read position     (std::get<1>) */
	const uint8_t* seed_reverse;    /* reverse flag      (std::get<2>) */
	int32_t initial_bandwidth;      /* -b */
	int32_t ramp_bandwidth;         /* -B (0 = none) */
} ga_batch;

typedef struct ga_read_result
{
	int32_t failed;                 /* AlignmentResult::alignmentFailed */
	int32_t score;                  /* alignment.score() (INT32_MAX when failed) */
	uint64_t alignment_start;       /* AlignmentResult::alignmentStart */
	uint64_t alignment_end;         /* AlignmentResult::alignmentEnd */
	int32_t query_position;         /* alignment.query_position() */
	uint32_t flags;                 /* GA_FLAG_* */
	uint64_t mapping_offset;        /* first entry in ga_results_mappings(); the reads' ranges need not follow each other */
	uint64_t n_mappings;            /* alignment.path().mapping_size() */
	uint64_t reserved;              /* EstimatedCorrectlyAligned of the chosen seed = 64 x retained slices (GraphAligner.h:371-374), also for a failed read */
	uint64_t n_trace;               /* AlignmentResult::trace.size(); the items come from ga_results_read_trace() */
	uint64_t word_columns;          /* forward-pass word updates spent on this read (all seeds, both directions) */
} ga_read_result;

#define GA_FLAG_STREAM_ERROR 1u     /* a DP stream hit a hard limit (e.g. band >= 200000 bp, reference alternate method) */
#define GA_FLAG_BAD_SEED 2u         /* seed node not in the graph / position outside the read (reference: std::out_of_range) */
#define GA_FLAG_BAD_CHAR 4u         /* read character the reference aborts on */
#define GA_FLAG_CYCLIC 8u           /* a band held a cyclic component */
#define GA_FLAG_RAMP_REDO 16u       /* -B: a stream went back and redid a stretch with the ramp bandwidth (GraphAligner.h:2648-2719) */
#define GA_FLAG_RAMP_STALE 32u      /* -B: after a redo the reference traced through a stretch re-computed from a sqrt checkpoint of the abandoned pass
                                       (GraphAligner.h:2667,2772-2786,2858-2943); the result is that one, as the reference reports it */

typedef struct ga_mapping            /* 32 bytes: a batch of 10 000 x 10 kbp reads on 32-bp nodes returns 3.3 million of them */
{
	int64_t node_id;                /* digraph node id (2*id / 2*id+1), as the reference's AlignOneWay returns it */
	uint32_t offset;
	uint32_t rank;
	int32_t from_length;            /* the single Edit of the mapping */
	int32_t to_length;
	uint32_t read_start;            /* edit.sequence == read.substr(read_start, to_length) */
	uint32_t is_reverse;
} ga_mapping;

typedef struct ga_trace_item        /* AlignmentResult::TraceItem */
{
	int32_t node_id;                /* bigraph node id */
	uint32_t offset;
	uint64_t readpos;
	uint8_t reverse;
	uint8_t type;                   /* 1 MATCH 2 MISMATCH 3 INSERTION 4 DELETION 5 FORWARDBACKWARDSPLIT */
	char graph_char;
	char read_char;
	uint32_t reserved;
} ga_trace_item;

/* host buffers in, host results out: H2D of the reads, kernels, traceback, D2H, result assembly.
 * The buffers behind `batch` are referenced, not copied: keep them valid until the results are freed. */
ga_results* ga_align_batch(ga_ctx* ctx, const ga_batch* batch);

/* the same in three steps, so that a caller can keep inputs resident in HBM and time the GPU part alone */
ga_staged* ga_stage_batch(ga_ctx* ctx, const ga_batch* batch);          /* split reads into DP streams, H2D */
int ga_run_staged(ga_ctx* ctx, ga_staged* staged);                      /* launch kernels (asynchronous) */
int ga_sync(ga_ctx* ctx);                                               /* wait for the context's stream */
ga_results* ga_finish_staged(ga_ctx* ctx, ga_staged* staged);           /* D2H + result assembly (may be called once per run) */
void ga_staged_free(ga_ctx* ctx, ga_staged* staged);
void* ga_cuda_stream(ga_ctx* ctx);                                      /* cudaStream_t the kernels run on */

size_t ga_results_count(const ga_results* r);
const ga_read_result* ga_results_reads(const ga_results* r);
const ga_mapping* ga_results_mappings(const ga_results* r);
/* AlignmentResult::trace of read i, materialised on demand from the device's compact move record: writes up to
 * `capacity` items and returns the read's item count (call with buffer == NULL to query it).  The ga_batch buffers the
 * results came from must still be valid. */
size_t ga_results_read_trace(const ga_results* r, size_t read_index, ga_trace_item* buffer, size_t capacity);
void ga_results_free(ga_results* r);
/* FNV-1a 64 over (node_id, offset, reverse, readpos, type) of read i's trace items, each as a little-endian u64:
 * a compact fingerprint of the whole path for differential tests */
uint64_t ga_results_trace_hash(const ga_results* r, size_t read_index);

typedef struct ga_stats
{
	uint64_t streams;               /* DP streams run (two per seed at most) */
	uint64_t word_columns;          /* forward-pass word updates, SURVEY 8d "W" */
	uint64_t retries;               /* streams re-run with larger scratch */
	uint64_t h2d_bytes;
	uint64_t d2h_bytes;
	uint64_t launches;              /* alignment kernel launches */
	uint64_t graph_bytes;           /* bytes of the graph replica on the device */
	/* device time of the three kernels of every launch sequence, CUDA events on the context's stream, microseconds */
	uint64_t peq_us;                /* the match-mask pre-pass and the bad-character check */
	uint64_t forward_us;            /* the bit-parallel DP (small-band or general forward kernel) */
	uint64_t trace_us;              /* the traceback */
} ga_stats;
int ga_get_stats(const ga_ctx* ctx, ga_stats* out);   /* cumulative since ga_create / ga_reset_stats */
int ga_reset_stats(ga_ctx* ctx);
/* measured INT32 lane-operations per second of a dependency-free LOP3/IADD3 kernel on this GPU: the denominator of
 * the integer-ALU roofline fraction.  Returns 0 on failure. */
double ga_measure_int32_peak(ga_ctx* ctx);

/* ---- a stream of batches: `depth` contexts of one GPU, one host thread each ---------------------------------
 * Replaces the reference's worker threads popping reads from a shared stack (Aligner.cpp:107-117,285-298): the unit
 * is a batch, and while batch i's kernel runs the host stages batch i+1 and assembles batch i-1.  Batches are served
 * round robin and results come back in submission order.  At most `depth` batches may be in flight: ga_pipeline_submit
 * returns -2 (and does nothing) when the pipeline is full.  The buffers behind a submitted batch must stay valid until
 * its results have been freed.  All calls on one pipeline come from one thread; the pipeline owns its worker threads. */
typedef struct ga_pipeline ga_pipeline;
ga_pipeline* ga_pipeline_create(int device, int depth);          /* depth 1..8 (2 = double buffering); NULL on failure */
void ga_pipeline_destroy(ga_pipeline* p);
const char* ga_pipeline_last_error(const ga_pipeline* p);
int ga_pipeline_depth(const ga_pipeline* p);
ga_ctx* ga_pipeline_context(ga_pipeline* p, int lane);           /* the lane's context (stats, stream); not for aligning */
int ga_pipeline_graph_upload(ga_pipeline* p, const ga_graph* g); /* one replica per lane's context */
int ga_pipeline_submit(ga_pipeline* p, const ga_batch* batch);   /* asynchronous; 0, -2 = full */
ga_results* ga_pipeline_next(ga_pipeline* p);                    /* oldest submitted batch's results (blocks); NULL on error */
int ga_pipeline_in_flight(const ga_pipeline* p);
int ga_pipeline_get_stats(ga_pipeline* p, ga_stats* out);        /* summed over the lanes; -1 while a lane is still working on a batch */
int ga_pipeline_reset_stats(ga_pipeline* p);

#ifdef __cplusplus
}
#endif
#endif
