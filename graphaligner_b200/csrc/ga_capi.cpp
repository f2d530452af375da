// extern "C" layer (include/graphaligner_b200.h) over the C++ host code.  Exceptions stop here.
#include "../../include/graphaligner_b200.h"
#include <cstring>
#include <exception>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
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
};

struct ga_results
{
	std::vector<ga_read_result> reads;
	std::vector<ga_mapping> mappings;
	std::vector<ga_trace_item> trace;
};

struct ga_staged
{
	std::vector<std::string> names;
	std::vector<std::string> sequences;
	std::vector<std::vector<ga::SeedHit>> seeds;
	std::vector<ga::ReadInput> reads;
	std::unique_ptr<ga::BatchPlan> plan;
	ga::StagedBatch* device = nullptr;
	int b = 0, B = 0;
};

static thread_local std::string g_globalError;

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
	const size_t n = batch->n_reads;
	st->names.resize(n);
	st->sequences.resize(n);
	st->seeds.resize(n);
	st->reads.resize(n);
	for (size_t i = 0; i < n; i++)
	{
		st->sequences[i].assign(batch->sequences + batch->seq_offsets[i], batch->sequences + batch->seq_offsets[i + 1]);
		if (batch->names && batch->name_offsets) st->names[i].assign(batch->names + batch->name_offsets[i], batch->names + batch->name_offsets[i + 1]);
		for (uint64_t k = batch->seed_offsets[i]; k < batch->seed_offsets[i + 1]; k++)
		{
			st->seeds[i].emplace_back((int)batch->seed_node[k], (size_t)batch->seed_pos[k], batch->seed_reverse[k] != 0);
		}
		st->reads[i] = ga::ReadInput { &st->names[i], &st->sequences[i], &st->seeds[i] };
	}
	st->b = batch->initial_bandwidth;
	st->B = batch->ramp_bandwidth;
}

static ga_results* packResults(const std::vector<ga::ReadInput>& reads, const std::vector<AlignmentResult>& results)
{
	ga_results* out = new ga_results();
	out->reads.resize(results.size());
	size_t nm = 0, nt = 0;
	for (auto& r : results)
	{
		if (!r.alignmentFailed) { nm += r.alignment.path.mapping.size(); nt += r.trace.size(); }
	}
	out->mappings.reserve(nm);
	out->trace.reserve(nt);
	for (size_t i = 0; i < results.size(); i++)
	{
		const AlignmentResult& r = results[i];
		ga_read_result& o = out->reads[i];
		memset(&o, 0, sizeof(o));
		o.failed = r.alignmentFailed ? 1 : 0;
		o.score = r.alignmentFailed ? std::numeric_limits<int32_t>::max() : r.alignment.score;
		o.flags = r.flags;
		o.word_columns = r.wordColumns;
		o.mapping_offset = out->mappings.size();
		o.trace_offset = out->trace.size();
		if (r.alignmentFailed) continue;
		o.alignment_start = r.alignmentStart;
		o.alignment_end = r.alignmentEnd;
		o.query_position = r.alignment.query_position;
		for (auto& m : r.alignment.path.mapping)
		{
			ga_mapping gm;
			memset(&gm, 0, sizeof(gm));
			gm.node_id = m.position.node_id;
			gm.offset = m.position.offset;
			gm.rank = m.rank;
			gm.is_reverse = m.position.is_reverse ? 1 : 0;
			if (!m.edit.empty())
			{
				gm.from_length = m.edit[0].from_length;
				gm.to_length = m.edit[0].to_length;
				gm.read_start = m.edit[0].read_start;
			}
			out->mappings.push_back(gm);
		}
		o.n_mappings = out->mappings.size() - o.mapping_offset;
		for (auto& t : r.trace)
		{
			ga_trace_item gt;
			memset(&gt, 0, sizeof(gt));
			gt.node_id = t.nodeID;
			gt.offset = (uint32_t)t.offset;
			gt.readpos = t.readpos;
			gt.reverse = t.reverse ? 1 : 0;
			gt.type = (uint8_t)t.type;
			gt.graph_char = t.graphChar;
			gt.read_char = t.readChar;
			out->trace.push_back(gt);
		}
		o.n_trace = out->trace.size() - o.trace_offset;
	}
	return out;
}

ga_staged* ga_stage_batch(ga_ctx* ctx, const ga_batch* batch)
{
	ga_staged* st = new ga_staged();
	int rc = guarded(ctx, [&]() {
		if (!ctx->graph) throw std::logic_error("ga_stage_batch: no graph uploaded");
		fillStaged(st, batch);
		st->plan.reset(new ga::BatchPlan(ctx->graph->graph, st->reads));
		st->device = ga::StageAndUpload(ctx->dev, st->plan->streams, st->plan->parts, st->b, st->B, &ctx->stats);
	});
	if (rc != 0)
	{
		delete st;
		return nullptr;
	}
	return st;
}

int ga_run_staged(ga_ctx* ctx, ga_staged* st)
{
	return guarded(ctx, [&]() { ga::RunStaged(ctx->dev, st->device); });
}

int ga_sync(ga_ctx* ctx)
{
	return guarded(ctx, [&]() { ga::SyncDevice(ctx->dev); });
}

ga_results* ga_finish_staged(ga_ctx* ctx, ga_staged* st)
{
	ga_results* res = nullptr;
	int rc = guarded(ctx, [&]() {
		std::vector<ga_stream_out> outs;
		std::vector<uint32_t> arena;
		ga::FinishStaged(ctx->dev, st->device, outs, arena, &ctx->stats);
		std::vector<AlignmentResult> results(st->reads.size());
		const AlignmentGraph& graph = ctx->graph->graph;
		for (size_t i = 0; i < st->reads.size(); i++)
		{
			if (st->reads[i].seeds->empty()) continue;   // stays failed: "has no seed hits" (Aligner.cpp:131-138)
			results[i] = ga::AssembleRead(graph, st->reads[i], *st->plan, (uint32_t)i, outs, arena);
		}
		for (size_t i = 0; i < results.size(); i++)
		{
			if (st->reads[i].seeds->empty()) results[i].alignment.score = std::numeric_limits<int32_t>::max();
		}
		ctx->stats.streams += st->plan->streams.size();
		for (auto& o : outs) ctx->stats.wordColumns += o.wordColumns;
		res = packResults(st->reads, results);
	});
	return rc == 0 ? res : nullptr;
}

void ga_staged_free(ga_ctx* ctx, ga_staged* st)
{
	if (!st) return;
	if (st->device) ga::FreeStaged(ctx->dev, st->device);
	delete st;
}

void* ga_cuda_stream(ga_ctx* ctx) { return ga::DeviceStream(ctx->dev); }

ga_results* ga_align_batch(ga_ctx* ctx, const ga_batch* batch)
{
	ga_staged* st = ga_stage_batch(ctx, batch);
	if (!st) return nullptr;
	ga_results* res = nullptr;
	if (ga_run_staged(ctx, st) == 0) res = ga_finish_staged(ctx, st);
	ga_staged_free(ctx, st);
	return res;
}

size_t ga_results_count(const ga_results* r) { return r->reads.size(); }
const ga_read_result* ga_results_reads(const ga_results* r) { return r->reads.data(); }
const ga_mapping* ga_results_mappings(const ga_results* r) { return r->mappings.data(); }
const ga_trace_item* ga_results_trace(const ga_results* r) { return r->trace.data(); }
void ga_results_free(ga_results* r) { delete r; }

uint64_t ga_results_trace_hash(const ga_results* r, size_t i)
{
	uint64_t h = 14695981039346656037ull;
	auto mix = [&h](uint64_t v) {
		for (int b = 0; b < 8; b++)
		{
			h ^= (v >> (8 * b)) & 0xff;
			h *= 1099511628211ull;
		}
	};
	const ga_read_result& rr = r->reads[i];
	for (uint64_t k = 0; k < rr.n_trace; k++)
	{
		const ga_trace_item& t = r->trace[rr.trace_offset + k];
		mix((uint64_t)(int64_t)t.node_id);
		mix(t.offset);
		mix(t.reverse);
		mix(t.readpos);
		mix(t.type);
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

}
