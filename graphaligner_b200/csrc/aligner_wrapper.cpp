#include "aligner_wrapper.h"
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <exception>
#include <limits>
#include <map>
#include <memory>
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
// key: graph identity (AlignmentGraph::Uid, not its address: a new graph may reuse the address of a destroyed one),
// device * 16 + lane (AlignReads streams a large read set through two contexts of every device)
std::map<std::pair<uint64_t, int>, Engine*> g_engines;

Engine* engineFor(const AlignmentGraph& graph, int device, int lane = 0)
{
	std::lock_guard<std::mutex> lock(g_enginesMutex);
	auto key = std::make_pair(graph.Uid(), device * 16 + lane);
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
		if (it->first.first == graph.Uid())
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
	return AlignReads(graph, reads, initialBandwidth, rampBandwidth, std::vector<int>(1, device));
}

namespace
{
// a batch that does not fit the device (history pool, scratch) is cut in half and tried again: the caller never loses the
// results of the other batches to one oversized one
bool isCapacityError(const std::exception& e)
{
	const std::string what = e.what();
	return what.find("out of memory") != std::string::npos || what.find("batch too large") != std::string::npos || what.find("allocating") != std::string::npos;
}

void alignRange(Engine* e, const AlignmentGraph& graph, const std::vector<ga::ReadInput>& inputs, size_t first, size_t last, int b, int B, std::mutex* gpuTurn, std::vector<AlignmentResult>& results)
{
	if (first >= last) return;
	std::vector<ga::ReadInput> part(inputs.begin() + first, inputs.begin() + last);
	try
	{
		std::vector<AlignmentResult> got = ga::AlignBatch(e->ctx, graph, part, b, B, nullptr, gpuTurn);
		for (size_t i = 0; i < got.size(); i++) results[first + i] = std::move(got[i]);
	}
	catch (const std::exception& ex)
	{
		if (last - first < 2 || !isCapacityError(ex)) throw;
		const size_t mid = first + (last - first) / 2;
		alignRange(e, graph, inputs, first, mid, b, B, gpuTurn, results);
		alignRange(e, graph, inputs, mid, last, b, B, gpuTurn, results);
	}
}
}

// The counterpart of the reference's worker pool (Aligner.cpp:107-117,285-306): the read set is cut into batches, every
// device runs two lanes (host threads with a context each) that pull batches from one queue, results land in input order.
std::vector<AlignmentResult> AlignReads(const AlignmentGraph& graph, const std::vector<AlignerRead>& reads, int initialBandwidth, int rampBandwidth, const std::vector<int>& devicesIn)
{
	std::vector<int> devices = devicesIn;
	if (devices.empty()) devices.push_back(0);
	std::vector<ga::ReadInput> inputs(reads.size());
	size_t totalBp = 0;
	for (size_t i = 0; i < reads.size(); i++)
	{
		inputs[i] = ga::ReadInput { reads[i].name.data(), reads[i].name.size(), reads[i].sequence.data(), reads[i].sequence.size(), reads[i].seedHits.data(), reads[i].seedHits.size() };
		totalBp += reads[i].sequence.size() * std::max<size_t>(1, reads[i].seedHits.size());
	}
	// A large read set goes through the GPUs as a stream of batches of ~GA_BATCH_BP read bases (default 100 Mbp, the size of
	// BASELINE config 2), two contexts deep per device: while one batch's kernel runs, the other lane's host thread plans and
	// uploads the next batch and assembles the previous one.  Reads are independent, so the cut points do not matter.
	size_t batchBp = 100000000;
	if (const char* e = getenv("GA_BATCH_BP")) batchBp = std::max<size_t>(1, (size_t)atoll(e));
	// several devices: at least two batches per device, so that all of them get work
	if (devices.size() > 1) batchBp = std::max<size_t>(1, std::min(batchBp, totalBp / (2 * devices.size()) + 1));
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
	std::vector<AlignmentResult> results(reads.size());
	if (nBatches <= 1)
	{
		Engine* e = engineFor(graph, devices[0]);
		std::lock_guard<std::mutex> lock(e->mutex);
		alignRange(e, graph, inputs, 0, reads.size(), initialBandwidth, rampBandwidth, nullptr, results);
		return results;
	}
	std::atomic<size_t> nextBatch(0);
	std::exception_ptr error;
	std::mutex errorMutex;
	// the two lanes of a device take turns on it; each one's planning and assembly run under the other's kernel
	std::vector<std::unique_ptr<std::mutex>> gpuTurn;
	for (size_t d = 0; d < devices.size(); d++) gpuTurn.emplace_back(new std::mutex());
	auto lane = [&](size_t deviceIndex, int laneIndex) {
		try
		{
			Engine* e = engineFor(graph, devices[deviceIndex], laneIndex);
			std::lock_guard<std::mutex> lock(e->mutex);
			while (true)
			{
				const size_t k = nextBatch.fetch_add(1);
				if (k >= nBatches) break;
				alignRange(e, graph, inputs, cuts[k], cuts[k + 1], initialBandwidth, rampBandwidth, gpuTurn[deviceIndex].get(), results);
			}
		}
		catch (...)
		{
			std::lock_guard<std::mutex> lock(errorMutex);
			if (!error) error = std::current_exception();
			nextBatch.store(nBatches);
		}
	};
	std::vector<std::thread> workers;
	for (size_t d = 0; d < devices.size(); d++)
	{
		for (int l = 0; l < 2; l++)
		{
			if (d == 0 && l == 0) continue;   // this thread
			workers.emplace_back(lane, d, l);
		}
	}
	lane(0, 0);
	for (auto& w : workers) w.join();
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
