#include "aligner_wrapper.h"
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <exception>
#include <limits>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include "ga_device.h"

namespace
{
struct Engine
{
	ga::DeviceCtx* ctx = nullptr;
	std::mutex mutex;   // a device context is not thread-safe; the reference's workers call AlignOneWay concurrently
};
std::mutex g_enginesMutex;
// key: graph, device * 16 + lane (AlignReads streams a large read set through two contexts of the device)
std::map<std::pair<const AlignmentGraph*, int>, Engine*> g_engines;

Engine* engineFor(const AlignmentGraph& graph, int device, int lane = 0)
{
	std::lock_guard<std::mutex> lock(g_enginesMutex);
	auto key = std::make_pair(&graph, device * 16 + lane);
	auto found = g_engines.find(key);
	if (found != g_engines.end()) return found->second;
	Engine* e = new Engine();
	e->ctx = ga::CreateDevice(device);
	ga::UploadGraph(e->ctx, graph);
	g_engines[key] = e;
	return e;
}
}

void ReleaseAlignerEngine(const AlignmentGraph& graph)
{
	std::lock_guard<std::mutex> lock(g_enginesMutex);
	for (auto it = g_engines.begin(); it != g_engines.end();)
	{
		if (it->first.first == &graph)
		{
			ga::DestroyDevice(it->second->ctx);
			delete it->second;
			it = g_engines.erase(it);
		}
		else ++it;
	}
}

std::vector<AlignmentResult> AlignReads(const AlignmentGraph& graph, const std::vector<AlignerRead>& reads, int initialBandwidth, int rampBandwidth, int device)
{
	std::vector<ga::ReadInput> inputs(reads.size());
	for (size_t i = 0; i < reads.size(); i++)
	{
		inputs[i] = ga::ReadInput { reads[i].name.data(), reads[i].name.size(), reads[i].sequence.data(), reads[i].sequence.size(), reads[i].seedHits.data(), reads[i].seedHits.size() };
	}
	// A large read set goes through the GPU as a stream of batches of ~GA_BATCH_BP read bases (default 100 Mbp, the size of
	// BASELINE config 2), two contexts deep: while one batch's kernel runs, the other lane's host thread plans, pads and
	// uploads the next batch and assembles the previous one.  Reads are independent, so the cut points do not matter.
	size_t batchBp = 100000000;
	if (const char* e = getenv("GA_BATCH_BP")) batchBp = std::max<size_t>(1, (size_t)atoll(e));
	std::vector<size_t> cuts(1, 0);
	size_t bp = 0;
	for (size_t i = 0; i < reads.size(); i++)
	{
		const size_t need = reads[i].sequence.size() * std::max<size_t>(1, reads[i].seedHits.size());
		if (bp + need > batchBp && i > cuts.back())
		{
			cuts.push_back(i);
			bp = 0;
		}
		bp += need;
	}
	cuts.push_back(reads.size());
	const size_t nBatches = cuts.size() - 1;
	if (nBatches <= 1)
	{
		Engine* e = engineFor(graph, device);
		std::lock_guard<std::mutex> lock(e->mutex);
		return ga::AlignBatch(e->ctx, graph, inputs, initialBandwidth, rampBandwidth, nullptr);
	}
	std::vector<AlignmentResult> results(reads.size());
	std::atomic<size_t> nextBatch(0);
	std::exception_ptr error;
	std::mutex errorMutex;
	std::mutex gpuTurn;   // the two lanes take turns on the device; each one's planning and assembly run under the other's kernel
	auto lane = [&](int laneIndex) {
		try
		{
			Engine* e = engineFor(graph, device, laneIndex);
			std::lock_guard<std::mutex> lock(e->mutex);
			while (true)
			{
				const size_t k = nextBatch.fetch_add(1);
				if (k >= nBatches) break;
				std::vector<ga::ReadInput> part(inputs.begin() + cuts[k], inputs.begin() + cuts[k + 1]);
				std::vector<AlignmentResult> got = ga::AlignBatch(e->ctx, graph, part, initialBandwidth, rampBandwidth, nullptr, &gpuTurn);
				for (size_t i = 0; i < got.size(); i++) results[cuts[k] + i] = std::move(got[i]);
			}
		}
		catch (...)
		{
			std::lock_guard<std::mutex> lock(errorMutex);
			if (!error) error = std::current_exception();
			nextBatch.store(nBatches);
		}
	};
	std::thread second(lane, 1);
	lane(0);
	second.join();
	if (error) std::rethrow_exception(error);
	return results;
}

AlignmentResult AlignOneWay(const AlignmentGraph& graph, const std::string& seq_id, const std::string& sequence, int initialBandwidth, int rampBandwidth, size_t dynamicRowStart,
	const std::vector<std::tuple<int, size_t, bool>>& seedHits)
{
	(void)dynamicRowStart;   // parsed but unused by the reference's algorithm as well
	auto t0 = std::chrono::system_clock::now();
	// the reference dereferences nodeLookup.at() and substr() and lets std::out_of_range escape (GraphAligner.h:423)
	for (auto& hit : seedHits)
	{
		if (!graph.HasNode(std::get<0>(hit) * 2)) throw std::out_of_range("AlignOneWay: seed node not in graph");
		if (std::get<1>(hit) >= sequence.size()) throw std::out_of_range("AlignOneWay: seed position outside the read");
	}
	std::vector<AlignerRead> one(1);
	one[0].name = seq_id;
	one[0].sequence = sequence;
	one[0].seedHits = seedHits;
	AlignmentResult r = std::move(AlignReads(graph, one, initialBandwidth, rampBandwidth)[0]);
	r.elapsedMilliseconds = (size_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now() - t0).count();
	return r;
}

AlignmentResult AlignOneWay(const AlignmentGraph& graph, const std::string& seq_id, const std::string& sequence, int initialBandwidth, int rampBandwidth, size_t dynamicRowStart)
{
	(void)graph; (void)seq_id; (void)sequence; (void)initialBandwidth; (void)rampBandwidth; (void)dynamicRowStart;
	AlignmentResult r;
	r.alignmentFailed = true;
	r.alignment.score = std::numeric_limits<int32_t>::max();
	return r;
}
