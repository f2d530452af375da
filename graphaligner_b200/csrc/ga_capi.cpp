// extern "C" layer (include/graphaligner_b200.h) over the C++ host code.  Exceptions stop here.
#include "../../include/graphaligner_b200.h"
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include "alignment_graph.h"
#include "bigraph_to_digraph.h"
#include "ga_device.h"
#include "ga_host.h"

struct ga_graph
{
	AlignmentGraph graph;
};

struct ga_ctx
{
	ga::DeviceCtx* dev = nullptr;
	const ga_graph* graph = nullptr;
	std::string error;
	ga::BatchStats stats;
	double budgetShare = 1.0;   // share of the device's free memory a batch of this context may plan with (ga_pipeline: 1 / depth)
	// ga_pipeline: the contexts of one GPU take turns on the kernel.  Two alignment kernels in flight at once share the SMs and
	// finish together, which puts the lanes in phase (both on the host, then both on the GPU) and loses the overlap.
	struct GpuTurn* gpuTurn = nullptr;
	uint64_t ticket = 0;       // ga_pipeline: the batch's place in the submission order
	bool ticketUsed = true;    // ... and whether its first launch has taken its place yet
};

// The lanes of a pipeline take the GPU one at a time, and for the first launch of every batch in the order the batches were
// submitted: results are handed out in that order, so an older batch waiting behind a younger one's kernels is pure delay.
struct GpuTurn
{
	std::mutex gpu;            // held while a lane's kernels run
	std::mutex m;
	std::condition_variable cv;
	uint64_t serving = 0;      // ticket whose first launch may go next
	void waitFor(uint64_t ticket)
	{
		std::unique_lock<std::mutex> lock(m);
		cv.wait(lock, [&]() { return serving == ticket; });
	}
	void next()
	{
		{
			std::lock_guard<std::mutex> lock(m);
			serving++;
		}
		cv.notify_all();
	}
};

struct ga_results
{
	ga::RawBuffer<ga_read_result> reads;
	// The mapping records.  Results of one launch sequence: they lie in the chunk's trace arena (mapBase points into it) - the
	// records the device wrote where it wrote them, the others in the room behind; ga_read_result::mapping_offset is relative to
	// mapBase either way.  Results merged from several launch sequences: copied into `mappings`, read after read.
	ga::RawBuffer<ga_mapping> mappings;
	const ga_mapping* mapBase = nullptr;
	// kept for the lazy trace items (ga_results_read_trace): what the device returned and what was decided per read.
	// A batch too large for the device is aligned in several launches (chunks of reads); each keeps its own buffers.
	const AlignmentGraph* graph = nullptr;
	std::vector<ga::ReadInput> inputs;
	struct Chunk
	{
		std::vector<ga_stream_in> streams;
		ga::RawBuffer<ga_stream_out> outs;
		ga::RawBuffer<uint32_t> arena;
	};
	std::vector<std::unique_ptr<Chunk>> chunks;
	struct Lazy
	{
		uint32_t chunk;
		int64_t fwStream, bwStream;
		size_t splitIndex;
		bool fwShifted;
		bool failed;
		size_t nTraceItems;
	};
	std::vector<Lazy> lazy;
};

struct ga_staged
{
	std::vector<ga::SeedHit> seeds;
	std::vector<ga::ReadInput> reads;
	std::unique_ptr<ga::BatchPlan> plan;
	ga::StagedBatch* device = nullptr;
	ga::DeviceCtx* dev = nullptr;   // the context it was staged on
	int b = 0, B = 0;
};

static thread_local std::string g_globalError;

// GA_TIMING=1: per-stage host timings on stderr
struct StageTimer
{
	bool on;
	std::chrono::steady_clock::time_point t;
	StageTimer() : on(getenv("GA_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
	void lap(const char* what)
	{
		if (!on) return;
		auto n = std::chrono::steady_clock::now();
		fprintf(stderr, "[ga timing] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
		t = n;
	}
};

// GA_TIMELINE=1: absolute time stamps (ms since the first one, with the context's address) of a batch's way through a lane
static void timeline(const void* ctx, const char* what)
{
	static const bool on = getenv("GA_TIMELINE") != nullptr;
	if (!on) return;
	static const auto t0 = std::chrono::steady_clock::now();
	fprintf(stderr, "[ga timeline] %9.2f ms  ctx %p  %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), ctx, what);
}

template <typename F>
static int guarded(ga_ctx* ctx, F&& f)
{
	try
	{
		f();
		return 0;
	}
	catch (const std::exception& e)
	{
		if (ctx) ctx->error = e.what(); else g_globalError = e.what();
	}
	catch (...)
	{
		if (ctx) ctx->error = "unknown error"; else g_globalError = "unknown error";
	}
	return -1;
}

extern "C" {

ga_ctx* ga_create(int device)
{
	ga_ctx* ctx = new ga_ctx();
	if (guarded(nullptr, [&]() { ctx->dev = ga::CreateDevice(device); }) != 0)
	{
		delete ctx;
		return nullptr;
	}
	return ctx;
}

void ga_destroy(ga_ctx* ctx)
{
	if (!ctx) return;
	guarded(nullptr, [&]() { ga::DestroyDevice(ctx->dev); });
	delete ctx;
}

const char* ga_last_error(const ga_ctx* ctx) { return ctx ? ctx->error.c_str() : ""; }
const char* ga_global_error(void) { return g_globalError.c_str(); }

ga_graph* ga_graph_new(void) { return new ga_graph(); }

int ga_graph_add_node(ga_graph* g, int32_t id, const char* sequence, size_t length, int reverse_node)
{
	return guarded(nullptr, [&]() { g->graph.AddNode(id, std::string(sequence, length), reverse_node != 0); });
}

int ga_graph_add_edge(ga_graph* g, int32_t from, int32_t to)
{
	return guarded(nullptr, [&]() { g->graph.AddEdgeNodeId(from, to); });
}

int ga_graph_set_dbg_overlap(ga_graph* g, int32_t overlap)
{
	g->graph.DBGOverlap = overlap;
	return 0;
}

int ga_graph_finalize(ga_graph* g)
{
	return guarded(nullptr, [&]() { g->graph.Finalize(64); });
}

ga_graph* ga_graph_from_bigraph(size_t n_nodes, const int64_t* ids, const char* sequences, const uint64_t* seq_offsets,
	size_t n_edges, const int64_t* from, const uint8_t* from_start, const int64_t* to, const uint8_t* to_end, int32_t gfa_overlap)
{
	ga_graph* g = nullptr;
	int rc = guarded(nullptr, [&]() {
		std::vector<DirectedGraph::BiNode> nodes(n_nodes);
		for (size_t i = 0; i < n_nodes; i++)
		{
			nodes[i].id = ids[i];
			nodes[i].sequence.assign(sequences + seq_offsets[i], sequences + seq_offsets[i + 1]);
		}
		std::vector<DirectedGraph::BiEdge> edges(n_edges);
		for (size_t i = 0; i < n_edges; i++) edges[i] = DirectedGraph::BiEdge { from[i], to[i], from_start[i] != 0, to_end[i] != 0 };
		g = new ga_graph { gfa_overlap < 0 ? DirectedGraph::BuildFromVG(nodes, edges) : DirectedGraph::BuildFromGFA(nodes, edges, gfa_overlap) };
	});
	return rc == 0 ? g : nullptr;
}

ga_graph* ga_graph_load_vg(const char* path)
{
	ga_graph* g = nullptr;
	int rc = guarded(nullptr, [&]() { g = new ga_graph { DirectedGraph::StreamVGGraphFromFile(path) }; });
	return rc == 0 ? g : nullptr;
}

ga_graph* ga_graph_load_gfa(const char* path)
{
	ga_graph* g = nullptr;
	int rc = guarded(nullptr, [&]() { g = new ga_graph { DirectedGraph::StreamGFAGraphFromFile(path) }; });
	return rc == 0 ? g : nullptr;
}

void ga_graph_free(ga_graph* g) { delete g; }
size_t ga_graph_node_count(const ga_graph* g) { return g->graph.NodeSize(); }
size_t ga_graph_size_bp(const ga_graph* g) { return g->graph.SizeInBp(); }
size_t ga_graph_edge_count(const ga_graph* g) { return g->graph.NumEdges(); }

int ga_graph_upload(ga_ctx* ctx, const ga_graph* g)
{
	return guarded(ctx, [&]() {
		ga::UploadGraph(ctx->dev, g->graph);
		ctx->graph = g;
	});
}

static void fillStaged(ga_staged* st, const ga_batch* batch)
{
	// nothing is copied: ReadInput points into the caller's buffers, which must stay valid until the results
	// (and any lazily requested trace items) are no longer needed
	const size_t n = batch->n_reads;
	// offsets are absolute into the seed arrays and need not start at 0 (a sub-batch is the same arrays, shifted)
	const uint64_t seedBase = n ? batch->seed_offsets[0] : 0;
	const uint64_t nSeeds = n ? batch->seed_offsets[n] - seedBase : 0;
	st->seeds.resize(nSeeds);
	for (uint64_t k = 0; k < nSeeds; k++) st->seeds[k] = ga::SeedHit((int)batch->seed_node[seedBase + k], (size_t)batch->seed_pos[seedBase + k], batch->seed_reverse[seedBase + k] != 0);
	st->reads.resize(n);
	for (size_t i = 0; i < n; i++)
	{
		ga::ReadInput& r = st->reads[i];
		r.seq = batch->sequences + batch->seq_offsets[i];
		r.seqLen = (size_t)(batch->seq_offsets[i + 1] - batch->seq_offsets[i]);
		if (batch->names && batch->name_offsets)
		{
			r.name = batch->names + batch->name_offsets[i];
			r.nameLen = (size_t)(batch->name_offsets[i + 1] - batch->name_offsets[i]);
		}
		else
		{
			r.name = "";
			r.nameLen = 0;
		}
		r.seeds = st->seeds.data() + (batch->seed_offsets[i] - seedBase);
		r.nSeeds = (size_t)(batch->seed_offsets[i + 1] - batch->seed_offsets[i]);
	}
	st->b = batch->initial_bandwidth;
	st->B = batch->ramp_bandwidth;
}

// the reads' records and, straight from the device's run records into their final place, the mappings
static_assert(sizeof(ga_mapping) == sizeof(GaDeviceMapping) && sizeof(ga_mapping) == GA_MAP_WORDS * sizeof(uint32_t), "the device writes ga_mapping records");

// mapTail: first word of the arena's room for host-written mapping records (ga::FinishStaged)
static void packResults(ga_results* out, const std::vector<ga::ReadAssembly>& as, const AlignmentGraph& graph, const std::vector<ga::ReadInput>& reads, const ga_stream_out* outs,
	uint32_t* arena, size_t arenaWords, size_t mapTail)
{
	const size_t n = as.size();
	out->reads.resize(n);
	out->lazy.resize(n);
	// device-written records stay where they are; the others are written into the room behind the device's part
	std::vector<uint64_t> mapOff(n + 1, 0);
	mapOff[0] = mapTail / GA_MAP_WORDS;
	for (size_t i = 0; i < n; i++) mapOff[i + 1] = mapOff[i] + ((as[i].failed || as[i].deviceMapped) ? 0 : as[i].nMappings);
	if (mapOff[n] * GA_MAP_WORDS > arenaWords) throw std::logic_error("packResults: the arena has no room for the mapping records");
	ga_mapping* const base = (ga_mapping*)arena;
	out->mapBase = base;
	ga::ParallelFor(n, [&](size_t i) {
		const ga::ReadAssembly& a = as[i];
		ga_read_result& o = out->reads.data()[i];
		memset(&o, 0, sizeof(o));
		o.failed = a.failed ? 1 : 0;
		o.score = a.failed ? std::numeric_limits<int32_t>::max() : a.score;
		o.flags = a.flags;
		o.word_columns = a.wordColumns;
		o.reserved = a.estimated;   // EstimatedCorrectlyAligned of the chosen seed (kept for the two-round seed loop, also for failed reads)
		o.mapping_offset = a.deviceMapped ? a.deviceMapWord / GA_MAP_WORDS : mapOff[i];
		out->lazy[i] = ga_results::Lazy { 0, a.fwStream, a.bwStream, a.splitIndex, a.fwShifted, a.failed, a.nTraceItems };
		if (a.failed) return;
		o.alignment_start = a.alignmentStart;
		o.alignment_end = a.alignmentEnd;
		o.query_position = a.queryPosition;
		o.n_mappings = a.nMappings;
		o.n_trace = a.nTraceItems;
		if (a.deviceMapped) return;
		ga_mapping* dst = base + mapOff[i];
		ga::WriteMappings(graph, reads[i], a, outs, arena, dst);
	});
}

static ga_staged* stageOn(ga_ctx* ctx, ga::DeviceCtx* dev, const ga_batch* batch)
{
	ga_staged* st = new ga_staged();
	int rc = guarded(ctx, [&]() {
		if (!ctx->graph) throw std::logic_error("ga_stage_batch: no graph uploaded");
		StageTimer tm;
		st->dev = dev;
		fillStaged(st, batch);
		tm.lap("stage: marshal reads");
		// reads that already lie in page-locked memory (cudaHostAlloc / cudaHostRegister by the caller) are uploaded from there;
		// otherwise they are copied into the context's pinned staging buffer first
		const bool inPlace = !st->reads.empty() && ga::IsPinnedHost(st->reads[0].seq) && getenv("GA_NO_INPLACE_INPUT") == nullptr;
		uint8_t* pinned = inPlace ? (uint8_t*)const_cast<char*>(st->reads[0].seq) : nullptr;
		st->plan.reset(new ga::BatchPlan(ctx->graph->graph, st->reads,
			[dev, &pinned, inPlace](size_t bytes) { if (inPlace) { ga::EnsureDeviceParts(dev, bytes); return pinned; } pinned = ga::AllocPinnedParts(dev, bytes); return pinned; },
			[dev, &pinned](size_t off, size_t bytes) { ga::UploadPartsRange(dev, pinned, off, bytes); }, inPlace));
		tm.lap("stage: plan + build parts");
		st->device = ga::StageAndUpload(dev, st->plan->streams, st->plan->parts, st->plan->partsBytes, st->b, st->B, &ctx->stats, true);
		ga::SetReadRanges(dev, st->device, st->plan->readOff);
		tm.lap("stage: layout + H2D");
	});
	if (rc != 0)
	{
		delete st;
		return nullptr;
	}
	return st;
}

ga_staged* ga_stage_batch(ga_ctx* ctx, const ga_batch* batch) { return stageOn(ctx, ctx->dev, batch); }

int ga_run_staged(ga_ctx* ctx, ga_staged* st)
{
	return guarded(ctx, [&]() { ga::RunStaged(st->dev, st->device); });
}

int ga_sync(ga_ctx* ctx)
{
	return guarded(ctx, [&]() { ga::SyncDevice(ctx->dev); });
}

ga_results* ga_finish_staged(ga_ctx* ctx, ga_staged* st)
{
	ga_results* res = new ga_results();
	int rc = guarded(ctx, [&]() {
		StageTimer tm;
		res->chunks.emplace_back(new ga_results::Chunk());
		ga_results::Chunk& ch = *res->chunks.back();
		size_t mapTail = 0;
		ga::FinishStaged(st->dev, st->device, ch.outs, ch.arena, &ctx->stats, &st->plan->badChar, &mapTail);
		tm.lap("finish: wait kernel + D2H");
		const AlignmentGraph& graph = ctx->graph->graph;
		const size_t n = st->reads.size();
		std::vector<ga::ReadAssembly> as(n);
		ga::ParallelFor(n, [&](size_t i) {
			if (st->reads[i].nSeeds == 0) return;   // stays failed: "has no seed hits" (Aligner.cpp:131-138)
			as[i] = ga::AssembleRead(graph, st->reads[i], *st->plan, (uint32_t)i, ch.outs.data(), ch.arena.data(), false);
		});
		tm.lap("finish: assemble reads");
		ctx->stats.streams += st->plan->streams.size();
		for (size_t i = 0; i < ch.outs.size(); i++) ctx->stats.wordColumns += ch.outs.data()[i].wordColumns;
		packResults(res, as, graph, st->reads, ch.outs.data(), ch.arena.data(), ch.arena.size(), mapTail);
		res->graph = &graph;
		res->inputs = st->reads;
		// the seed lists live in the staged batch, which does not outlive this call; nothing reads them from the results
		for (ga::ReadInput& in : res->inputs) { in.seeds = nullptr; in.nSeeds = 0; }
		ch.streams = st->plan->streams;
		tm.lap("finish: pack results");
	});
	if (rc != 0)
	{
		delete res;
		return nullptr;
	}
	return res;
}

void ga_staged_free(ga_ctx* ctx, ga_staged* st)
{
	if (!st) return;
	if (st->device) ga::FreeStaged(st->dev, st->device);
	delete st;
}

void* ga_cuda_stream(ga_ctx* ctx) { return ga::DeviceStream(ctx->dev); }

// Combines the results of consecutive sub-batches (part k holds reads [cuts[k], cuts[k + 1])) into one result set:
// one allocation for all mappings, parts copied in parallel.
static ga_results* mergeParts(ga_ctx* ctx, std::vector<ga_results*>& parts, const std::vector<size_t>& cuts)
{
	ga_results* all = new ga_results();
	const size_t n = cuts.back();
	all->reads.resize(n);
	all->lazy.resize(n);
	all->inputs.resize(n);
	all->graph = &ctx->graph->graph;
	// the parts' records lie in their own arenas: copied read after read into one array
	std::vector<uint64_t> newOff(n + 1, 0);
	for (size_t k = 0; k < parts.size(); k++)
	{
		for (size_t i = 0; i < parts[k]->lazy.size(); i++) newOff[cuts[k] + i + 1] = parts[k]->reads.data()[i].failed ? 0 : parts[k]->reads.data()[i].n_mappings;
	}
	for (size_t i = 0; i < n; i++) newOff[i + 1] += newOff[i];
	all->mappings.resize(newOff[n]);
	all->mapBase = all->mappings.data();
	for (size_t k = 0; k < parts.size(); k++)
	{
		const ga_results* part = parts[k];
		const size_t first = cuts[k];
		ga::ParallelFor(part->lazy.size(), [&](size_t i) {
			const ga_read_result& r = part->reads.data()[i];
			if (r.failed || r.n_mappings == 0) return;
			memcpy(all->mappings.data() + newOff[first + i], part->mapBase + r.mapping_offset, (size_t)r.n_mappings * sizeof(ga_mapping));
		});
	}
	for (size_t k = 0; k < parts.size(); k++)
	{
		ga_results* part = parts[k];
		const uint32_t chunkIndex = (uint32_t)all->chunks.size();
		const size_t first = cuts[k];
		for (size_t i = 0; i < part->lazy.size(); i++)
		{
			all->reads.data()[first + i] = part->reads.data()[i];
			all->reads.data()[first + i].mapping_offset = newOff[first + i];
			all->lazy[first + i] = part->lazy[i];
			all->lazy[first + i].chunk = chunkIndex;
			all->inputs[first + i] = part->inputs[i];
		}
		all->chunks.push_back(std::move(part->chunks[0]));
		delete part;
		parts[k] = nullptr;
	}
	return all;
}

// a sub-batch is the same arrays with shifted offsets pointers: offsets are absolute, so only the bases move
static ga_batch subBatch(const ga_batch* batch, size_t first, size_t last)
{
	ga_batch sub = *batch;
	sub.n_reads = last - first;
	sub.seq_offsets = batch->seq_offsets + first;
	sub.name_offsets = batch->name_offsets ? batch->name_offsets + first : nullptr;
	sub.seed_offsets = batch->seed_offsets + first;
	return sub;
}

// kernels, then D2H + assembly; inside a pipeline the kernel part runs under the GPU's turn lock (the staging before and
// the D2H and host work after overlap with the other lanes' kernels)
static ga_results* runAndFinish(ga_ctx* ctx, ga_staged* st)
{
	if (!ctx->gpuTurn) return ga_run_staged(ctx, st) == 0 ? ga_finish_staged(ctx, st) : nullptr;
	{
		// the uploads of this batch finish under the other lanes' kernels, not inside this lane's turn
		if (ga_sync(ctx) != 0) return nullptr;
		timeline(ctx, "staged and uploaded, waiting for the GPU's turn");
		if (!ctx->ticketUsed) ctx->gpuTurn->waitFor(ctx->ticket);
		std::lock_guard<std::mutex> turn(ctx->gpuTurn->gpu);
		if (!ctx->ticketUsed)
		{
			ctx->ticketUsed = true;
			ctx->gpuTurn->next();   // the next batch may queue up behind this lock
		}
		timeline(ctx, "turn taken, kernels launched");
		if (ga_run_staged(ctx, st) != 0 || ga_sync(ctx) != 0) return nullptr;
		timeline(ctx, "kernels done, turn released");
	}
	ga_results* r = ga_finish_staged(ctx, st);
	timeline(ctx, "finished (D2H, assembly, packing)");
	return r;
}

static ga_results* alignAllSeeds(ga_ctx* ctx, const ga_batch* batch);

// ---- seeds in two rounds -------------------------------------------------------------------------------------------
// The reference takes a read's seeds one after the other and skips a seed whose (node, read position) lies on the trace of an
// earlier seed's alignment (GraphAligner.h:420-450).  With seeds from one true locus - the usual output of a seeder - every
// seed after the first is skipped.  Aligning all seeds at once (and replaying the rule afterwards) therefore does several
// times the reference's DP work.  Two rounds: the first seed of every read; then, for the reads that have more, the seeds the
// first alignment does not cover (all of them at once, the rule replayed among them).  The outcome is the reference's: a seed
// is run unless the first seed's trace covers it, the replay inside round two applies the traces of round two, and the first
// seed keeps the read unless a later one has a strictly larger EstimatedCorrectlyAligned (GraphAligner.h:434-449).
static ga_results* combineRounds(ga_ctx* ctx, ga_results* r1, ga_results* r2, size_t n)
{
	ga_results* all = new ga_results();
	all->reads.resize(n);
	all->lazy.resize(n);
	all->inputs = r1->inputs;
	all->graph = r1->graph;
	const uint32_t chunks1 = (uint32_t)r1->chunks.size();
	// which round decides a read: round two when it ran seeds of the read and either hit a stream error there (the loop ends with
	// a failure, as it would have with all seeds at once) or found a strictly larger estimate
	std::vector<uint8_t> fromTwo(n, 0);
	std::vector<uint64_t> newOff(n + 1, 0);
	for (size_t i = 0; i < n; i++)
	{
		const ga_read_result& a = r1->reads.data()[i];
		const ga_read_result& b = r2->reads.data()[i];
		const bool ranTwo = b.word_columns > 0 || b.flags != 0 || !b.failed;
		fromTwo[i] = ranTwo && ((b.flags & GA_FLAG_STREAM_ERROR) || b.reserved > a.reserved) ? 1 : 0;
		const ga_read_result& pick = fromTwo[i] ? b : a;
		newOff[i + 1] = newOff[i] + (pick.failed ? 0 : pick.n_mappings);
	}
	all->mappings.resize(newOff[n]);
	all->mapBase = all->mappings.data();
	ga::ParallelFor(n, [&](size_t i) {
		const ga_results* part = fromTwo[i] ? r2 : r1;
		ga_read_result o = part->reads.data()[i];
		if (!o.failed && o.n_mappings) memcpy(all->mappings.data() + newOff[i], part->mapBase + o.mapping_offset, (size_t)o.n_mappings * sizeof(ga_mapping));
		o.mapping_offset = newOff[i];
		o.flags = r1->reads.data()[i].flags | r2->reads.data()[i].flags;
		o.word_columns = r1->reads.data()[i].word_columns + r2->reads.data()[i].word_columns;
		all->reads.data()[i] = o;
		all->lazy[i] = part->lazy[i];
		if (fromTwo[i]) all->lazy[i].chunk += chunks1;
	});
	for (auto& c : r1->chunks) all->chunks.push_back(std::move(c));
	for (auto& c : r2->chunks) all->chunks.push_back(std::move(c));
	delete r1;
	delete r2;
	(void)ctx;
	return all;
}

ga_results* ga_align_batch(ga_ctx* ctx, const ga_batch* batch)
{
	const size_t n = batch->n_reads;
	static const bool oneRound = getenv("GA_ALL_SEEDS_AT_ONCE") != nullptr;   // A/B measurements
	bool several = false;
	for (size_t i = 0; i < n && !several && !oneRound; i++) several = batch->seed_offsets[i + 1] - batch->seed_offsets[i] > 1;
	if (!several || !ctx->graph) return alignAllSeeds(ctx, batch);
	const AlignmentGraph& graph = ctx->graph->graph;
	// round one: the first seed of every read; a read with a seed the reference would throw on keeps all its seeds here (the
	// failure is then reported exactly as before)
	std::vector<uint64_t> off1(n + 1, 0), off2(n + 1, 0);
	std::vector<int32_t> node1, node2;
	std::vector<uint64_t> pos1, pos2;
	std::vector<uint8_t> rev1, rev2;
	std::vector<uint8_t> twoRounds(n, 0);
	auto push = [&](std::vector<int32_t>& nd, std::vector<uint64_t>& ps, std::vector<uint8_t>& rv, uint64_t k) {
		nd.push_back(batch->seed_node[k]); ps.push_back(batch->seed_pos[k]); rv.push_back(batch->seed_reverse[k]);
	};
	for (size_t i = 0; i < n; i++)
	{
		const uint64_t first = batch->seed_offsets[i], last = batch->seed_offsets[i + 1];
		bool valid = true;
		if (last - first > 1)
		{
			ga::ReadInput r;
			r.name = ""; r.nameLen = 0; r.seeds = nullptr; r.nSeeds = 0;
			r.seq = batch->sequences + batch->seq_offsets[i];
			r.seqLen = (size_t)(batch->seq_offsets[i + 1] - batch->seq_offsets[i]);
			for (uint64_t k = first; k < last && valid; k++) valid = ga::SeedIsValid(graph, r, ga::SeedHit((int)batch->seed_node[k], (size_t)batch->seed_pos[k], batch->seed_reverse[k] != 0));
		}
		twoRounds[i] = last - first > 1 && valid ? 1 : 0;
		for (uint64_t k = first; k < (twoRounds[i] ? first + 1 : last); k++) push(node1, pos1, rev1, k);
		off1[i + 1] = node1.size();
	}
	if (node1.empty()) { node1.push_back(0); pos1.push_back(0); rev1.push_back(0); }
	ga_batch b1 = *batch;
	b1.seed_offsets = off1.data(); b1.seed_node = node1.data(); b1.seed_pos = pos1.data(); b1.seed_reverse = rev1.data();
	ga_results* r1 = alignAllSeeds(ctx, &b1);
	if (!r1) return nullptr;
	// round two: the other seeds of those reads, unless the first alignment covers them or ended the read with a stream error
	int rc = guarded(ctx, [&]() {
		std::vector<std::tuple<size_t, size_t, size_t>> tried;
		for (size_t i = 0; i < n; i++)
		{
			if (twoRounds[i] && !(r1->reads.data()[i].flags & (GA_FLAG_STREAM_ERROR | GA_FLAG_BAD_CHAR)))
			{
				const ga_results::Lazy& lz = r1->lazy[i];
				const ga_results::Chunk& ch = *r1->chunks[lz.chunk];
				tried.clear();
				ga::CollectTried(graph, ch.outs.data(), ch.arena.data(), lz.fwStream, lz.bwStream, lz.splitIndex, lz.fwShifted, tried);
				for (uint64_t k = batch->seed_offsets[i] + 1; k < batch->seed_offsets[i + 1]; k++)
				{
					if (!ga::SeedCovered(graph, tried, ga::SeedHit((int)batch->seed_node[k], (size_t)batch->seed_pos[k], batch->seed_reverse[k] != 0))) push(node2, pos2, rev2, k);
				}
			}
			off2[i + 1] = node2.size();
		}
	});
	if (rc != 0) { delete r1; return nullptr; }
	if (node2.empty()) return r1;
	ga_batch b2 = *batch;
	b2.seed_offsets = off2.data(); b2.seed_node = node2.data(); b2.seed_pos = pos2.data(); b2.seed_reverse = rev2.data();
	ga_results* r2 = alignAllSeeds(ctx, &b2);
	if (!r2) { delete r1; return nullptr; }
	ga_results* all = nullptr;
	rc = guarded(ctx, [&]() { all = combineRounds(ctx, r1, r2, n); });
	return rc == 0 ? all : nullptr;
}

static ga_results* alignAllSeeds(ga_ctx* ctx, const ga_batch* batch)
{
	// split the batch when its DP history would not fit the device (the history is ~64 B per band column and slice)
	std::vector<size_t> cuts;   // chunk boundaries in reads
	StageTimer tmAll;
	int rc = guarded(ctx, [&]() {
		if (!ctx->graph) throw std::logic_error("ga_align_batch: no graph uploaded");
		const size_t budget = (size_t)(ga::FreeDeviceBytes(ctx->dev) * 0.8 * ctx->budgetShare);
		const int bw = std::max(batch->initial_bandwidth, batch->ramp_bandwidth);
		size_t used = 0;
		cuts.push_back(0);
		for (size_t i = 0; i < batch->n_reads; i++)
		{
			const size_t len = (size_t)(batch->seq_offsets[i + 1] - batch->seq_offsets[i]);
			const size_t nSeeds = (size_t)(batch->seed_offsets[i + 1] - batch->seed_offsets[i]);
			const size_t need = nSeeds * ga::EstimateStreamBytes(ctx->dev, len + 64, bw);
			if (used + need > budget && i > cuts.back())
			{
				cuts.push_back(i);
				used = 0;
			}
			used += need;
		}
		cuts.push_back(batch->n_reads);
	});
	if (rc != 0) return nullptr;
	tmAll.lap("align: split plan");
	std::vector<ga_results*> parts;
	auto fail = [&]() -> ga_results* {
		for (ga_results* p : parts) delete p;
		return nullptr;
	};
	// A chunk the device cannot hold after all (the estimate is an estimate: long-node graphs, wide bands) is cut in half
	// and tried again - the results of the other chunks are kept.
	// (Running two half-batches through two sets of device buffers and streams, to overlap staging and assembly with the
	// kernel, was measured: each half's kernel takes as long as the whole batch's - a stream's time is a chain of
	// latencies, not a share of the GPU - so nothing is gained.)
	auto capacityError = [&]() {
		const std::string& e = ctx->error;
		return e.find("out of memory") != std::string::npos || e.find("allocating") != std::string::npos || e.find("batch too large") != std::string::npos;
	};
	std::vector<size_t> done(1, 0);   // boundaries of the chunks as they were finally run
	std::vector<std::pair<size_t, size_t>> todo;
	for (size_t c = cuts.size() - 1; c-- > 0;) todo.emplace_back(cuts[c], cuts[c + 1]);   // a stack: first chunk on top
	while (!todo.empty())
	{
		const std::pair<size_t, size_t> range = todo.back();
		todo.pop_back();
		ga_batch sub = subBatch(batch, range.first, range.second);
		ga_staged* st = ga_stage_batch(ctx, &sub);
		ga_results* part = st ? runAndFinish(ctx, st) : nullptr;
		if (st) ga_staged_free(ctx, st);
		if (!part)
		{
			if (range.second - range.first < 2 || !capacityError()) return fail();
			const size_t mid = range.first + (range.second - range.first) / 2;
			todo.emplace_back(mid, range.second);
			todo.emplace_back(range.first, mid);
			continue;
		}
		parts.push_back(part);
		done.push_back(range.second);
	}
	if (parts.size() == 1)
	{
		ga_results* only = parts[0];
		parts.clear();
		return only;
	}
	ga_results* all = nullptr;
	rc = guarded(ctx, [&]() { all = mergeParts(ctx, parts, done); });
	return rc == 0 ? all : fail();
}

size_t ga_results_count(const ga_results* r) { return r->reads.size(); }
const ga_read_result* ga_results_reads(const ga_results* r) { return r->reads.data(); }
const ga_mapping* ga_results_mappings(const ga_results* r) { return r->mapBase ? r->mapBase : r->mappings.data(); }
void ga_results_free(ga_results* r)
{
	StageTimer tm;
	delete r;
	tm.lap("results: free");
}

static void materializeTrace(const ga_results* r, size_t i, std::vector<AlignmentResult::TraceItem>& items)
{
	const ga_results::Lazy& lz = r->lazy[i];
	ga::ReadAssembly as;
	as.failed = lz.failed;
	as.fwStream = lz.fwStream;
	as.bwStream = lz.bwStream;
	as.splitIndex = lz.splitIndex;
	as.fwShifted = lz.fwShifted;
	as.nTraceItems = lz.nTraceItems;
	const ga_results::Chunk& ch = *r->chunks[lz.chunk];
	ga::BuildTraceItems(*r->graph, r->inputs[i], as, ch.streams.data(), ch.outs.data(), ch.arena.data(), items);
}

size_t ga_results_read_trace(const ga_results* r, size_t i, ga_trace_item* buffer, size_t capacity)
{
	if (i >= r->lazy.size() || r->lazy[i].failed) return 0;
	if (buffer == nullptr || capacity < r->lazy[i].nTraceItems) return r->lazy[i].nTraceItems;
	std::vector<AlignmentResult::TraceItem> items;
	try
	{
		materializeTrace(r, i, items);
	}
	catch (...)
	{
		return 0;
	}
	for (size_t k = 0; k < items.size() && k < capacity; k++)
	{
		const AlignmentResult::TraceItem& t = items[k];
		ga_trace_item gt;
		memset(&gt, 0, sizeof(gt));
		gt.node_id = t.nodeID;
		gt.offset = (uint32_t)t.offset;
		gt.readpos = t.readpos;
		gt.reverse = t.reverse ? 1 : 0;
		gt.type = (uint8_t)t.type;
		gt.graph_char = t.graphChar;
		gt.read_char = t.readChar;
		buffer[k] = gt;
	}
	return items.size();
}

uint64_t ga_results_trace_hash(const ga_results* r, size_t i)
{
	uint64_t h = 14695981039346656037ull;
	if (i >= r->lazy.size() || r->lazy[i].failed) return h;
	auto mix = [&h](uint64_t v) {
		for (int b = 0; b < 8; b++)
		{
			h ^= (v >> (8 * b)) & 0xff;
			h *= 1099511628211ull;
		}
	};
	std::vector<AlignmentResult::TraceItem> items;
	try
	{
		materializeTrace(r, i, items);
	}
	catch (...)
	{
		return 0;
	}
	for (auto& t : items)
	{
		mix((uint64_t)(int64_t)t.nodeID);
		mix(t.offset);
		mix(t.reverse ? 1 : 0);
		mix(t.readpos);
		mix((uint64_t)t.type);
	}
	return h;
}

int ga_get_stats(const ga_ctx* ctx, ga_stats* out)
{
	out->streams = ctx->stats.streams;
	out->word_columns = ctx->stats.wordColumns;
	out->retries = ctx->stats.retries;
	out->h2d_bytes = ctx->stats.h2dBytes;
	out->d2h_bytes = ctx->stats.d2hBytes;
	out->launches = ctx->stats.launches;
	out->graph_bytes = ga::GraphBytesOnDevice(ctx->dev);
	out->peq_us = (uint64_t)(ctx->stats.peqMs * 1000.0);
	out->forward_us = (uint64_t)(ctx->stats.forwardMs * 1000.0);
	out->trace_us = (uint64_t)(ctx->stats.traceMs * 1000.0);
	return 0;
}

double ga_measure_int32_peak(ga_ctx* ctx)
{
	double v = 0;
	guarded(ctx, [&]() { v = ga::MeasureInt32Peak(ctx->dev); });
	return v;
}

int ga_reset_stats(ga_ctx* ctx)
{
	ctx->stats = ga::BatchStats();
	return 0;
}

// ---- a stream of batches through `depth` contexts of one GPU -------------------------------------------------------
// The reference keeps its cores busy with N worker threads popping reads from one stack (Aligner.cpp:107-117,285-298).
// Here the unit is a batch and the resource is the GPU: while the kernel of batch i runs, the host plans, pads and
// uploads batch i+1 and assembles batch i-1.  Each lane = one context (own stream, device pools, pinned staging) driven
// by one host thread; batches go to the lanes round robin and come back in submission order.
struct ga_pipeline
{
	struct Lane
	{
		ga_ctx* ctx = nullptr;
		std::thread worker;
		ga_batch batch;
		uint64_t ticket = 0;
		ga_results* result = nullptr;
		bool busy = false;    // a batch was submitted and its result not taken yet
		bool done = false;    // ... and the worker has finished it
		std::string error;
	};
	std::vector<std::unique_ptr<Lane>> lanes;
	std::mutex m;
	GpuTurn gpuTurn;      // see ga_ctx::gpuTurn
	std::condition_variable cv;
	uint64_t submitted = 0, taken = 0;
	bool stop = false;
	std::string error;
};

static void pipelineWorker(ga_pipeline* p, ga_pipeline::Lane* lane)
{
	std::unique_lock<std::mutex> lock(p->m);
	while (true)
	{
		p->cv.wait(lock, [&]() { return p->stop || (lane->busy && !lane->done); });
		if (p->stop) return;
		lock.unlock();
		timeline(lane->ctx, "batch picked up by the lane");
		lane->ctx->ticket = lane->ticket;
		lane->ctx->ticketUsed = lane->ctx->gpuTurn == nullptr;
		ga_results* r = ga_align_batch(lane->ctx, &lane->batch);
		if (!lane->ctx->ticketUsed)
		{
			// the batch never reached a launch (an error, no reads): its place in the order is given up, in order
			lane->ctx->gpuTurn->waitFor(lane->ticket);
			lane->ctx->gpuTurn->next();
			lane->ctx->ticketUsed = true;
		}
		lock.lock();
		lane->result = r;
		if (!r) lane->error = lane->ctx->error;
		lane->done = true;
		p->cv.notify_all();
	}
}

ga_pipeline* ga_pipeline_create(int device, int depth)
{
	if (depth < 1 || depth > 8)
	{
		g_globalError = "ga_pipeline_create: depth must be 1..8";
		return nullptr;
	}
	ga_pipeline* p = new ga_pipeline();
	for (int i = 0; i < depth; i++)
	{
		ga_ctx* ctx = ga_create(device);
		if (!ctx)
		{
			for (auto& l : p->lanes) ga_destroy(l->ctx);
			delete p;
			return nullptr;
		}
		ctx->budgetShare = 1.0 / depth;
		if (depth > 1 && !getenv("GA_PIPELINE_FREE_RUN")) ctx->gpuTurn = &p->gpuTurn;   // env: A/B measurements without the turn lock
		p->lanes.emplace_back(new ga_pipeline::Lane());
		p->lanes.back()->ctx = ctx;
	}
	for (auto& l : p->lanes) l->worker = std::thread(pipelineWorker, p, l.get());
	return p;
}

void ga_pipeline_destroy(ga_pipeline* p)
{
	if (!p) return;
	{
		std::lock_guard<std::mutex> lock(p->m);
		p->stop = true;
	}
	p->cv.notify_all();
	for (auto& l : p->lanes)
	{
		if (l->worker.joinable()) l->worker.join();
		if (l->result) ga_results_free(l->result);
		ga_destroy(l->ctx);
	}
	delete p;
}

const char* ga_pipeline_last_error(const ga_pipeline* p) { return p ? p->error.c_str() : ""; }
int ga_pipeline_depth(const ga_pipeline* p) { return p ? (int)p->lanes.size() : 0; }
ga_ctx* ga_pipeline_context(ga_pipeline* p, int lane) { return (p && lane >= 0 && (size_t)lane < p->lanes.size()) ? p->lanes[lane]->ctx : nullptr; }

int ga_pipeline_graph_upload(ga_pipeline* p, const ga_graph* g)
{
	std::lock_guard<std::mutex> lock(p->m);
	if (p->submitted != p->taken)
	{
		p->error = "ga_pipeline_graph_upload: batches in flight";
		return -1;
	}
	for (auto& l : p->lanes)
	{
		if (ga_graph_upload(l->ctx, g) != 0)
		{
			p->error = l->ctx->error;
			return -1;
		}
	}
	return 0;
}

int ga_pipeline_submit(ga_pipeline* p, const ga_batch* batch)
{
	std::lock_guard<std::mutex> lock(p->m);
	ga_pipeline::Lane& lane = *p->lanes[p->submitted % p->lanes.size()];
	if (lane.busy)
	{
		p->error = "ga_pipeline_submit: pipeline full, take a result with ga_pipeline_next first";
		return -2;
	}
	lane.batch = *batch;
	lane.ticket = p->submitted;
	lane.result = nullptr;
	lane.error.clear();
	lane.done = false;
	lane.busy = true;
	p->submitted++;
	p->cv.notify_all();
	return 0;
}

ga_results* ga_pipeline_next(ga_pipeline* p)
{
	std::unique_lock<std::mutex> lock(p->m);
	if (p->taken == p->submitted)
	{
		p->error = "ga_pipeline_next: nothing in flight";
		return nullptr;
	}
	ga_pipeline::Lane& lane = *p->lanes[p->taken % p->lanes.size()];
	p->cv.wait(lock, [&]() { return lane.done; });
	ga_results* r = lane.result;
	if (!r) p->error = lane.error;
	lane.result = nullptr;
	lane.busy = false;
	lane.done = false;
	p->taken++;
	return r;
}

int ga_pipeline_in_flight(const ga_pipeline* p) { return p ? (int)(p->submitted - p->taken) : 0; }

int ga_pipeline_get_stats(ga_pipeline* p, ga_stats* out)
{
	std::lock_guard<std::mutex> lock(p->m);
	memset(out, 0, sizeof(*out));
	// the lanes' workers update their contexts' counters while a batch is in flight: a snapshot is only taken of idle lanes
	for (auto& l : p->lanes)
	{
		if (l->busy && !l->done)
		{
			p->error = "ga_pipeline_get_stats: batches in flight (take their results first)";
			return -1;
		}
	}
	for (auto& l : p->lanes)
	{
		ga_stats s;
		ga_get_stats(l->ctx, &s);
		out->streams += s.streams;
		out->word_columns += s.word_columns;
		out->retries += s.retries;
		out->h2d_bytes += s.h2d_bytes;
		out->d2h_bytes += s.d2h_bytes;
		out->launches += s.launches;
		out->graph_bytes += s.graph_bytes;
		out->peq_us += s.peq_us;
		out->forward_us += s.forward_us;
		out->trace_us += s.trace_us;
	}
	return 0;
}

int ga_pipeline_reset_stats(ga_pipeline* p)
{
	std::lock_guard<std::mutex> lock(p->m);
	if (p->submitted != p->taken)
	{
		p->error = "ga_pipeline_reset_stats: batches in flight";
		return -1;
	}
	for (auto& l : p->lanes) ga_reset_stats(l->ctx);
	return 0;
}

}
