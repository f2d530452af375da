#include "ga_host.h"
#include "../../include/graphaligner_b200.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <condition_variable>
#include <mutex>
#include <thread>

namespace ga
{

static const uint32_t FLAG_STREAM_ERROR = 1;   // a stream hit a hard limit (see ga_stream_out.status)
static const uint32_t FLAG_BAD_SEED = 2;       // seed node not in graph / position outside read
static const uint32_t FLAG_BAD_CHAR = 4;       // read holds a character the reference aborts on
static const uint32_t FLAG_CYCLIC = 8;         // some band held a cyclic component
static const uint32_t FLAG_RAMP_REDO = 16;     // -B: a stream went back and redid a stretch with the wide band
static const uint32_t FLAG_RAMP_STALE = 32;    // -B: ... and the reference's stale sqrt checkpoint decided part of the trace (GA_RAMP_STALE_BIT of rampRedos)

std::string ReverseComplement(const std::string& str)
{
	std::string result;
	result.reserve(str.size());
	for (size_t k = str.size(); k-- > 0;)
	{
		char c = str[k];
		switch (c)
		{
			case 'A': case 'a': result += 'T'; break;
			case 'C': case 'c': result += 'G'; break;
			case 'T': case 't': result += 'A'; break;
			case 'G': case 'g': result += 'C'; break;
			case 'N': case 'n': result += 'N'; break;
			case 'U': case 'u': result += 'A'; break;
			case 'R': case 'r': result += 'Y'; break;
			case 'Y': case 'y': result += 'R'; break;
			case 'K': case 'k': result += 'M'; break;
			case 'M': case 'm': result += 'K'; break;
			case 'S': case 's': result += 'S'; break;
			case 'W': case 'w': result += 'W'; break;
			case 'B': case 'b': result += 'V'; break;
			case 'V': case 'v': result += 'B'; break;
			case 'D': case 'd': result += 'H'; break;
			case 'H': case 'h': result += 'D'; break;
			default: break;   // reference: assert(false) -> character dropped under NDEBUG
		}
	}
	return result;
}

bool ValidReadChar(char c)
{
	switch (c)
	{
		case 'A': case 'a': case 'C': case 'c': case 'G': case 'g': case 'T': case 't': case 'N': case 'n':
		case 'R': case 'r': case 'Y': case 'y': case 'K': case 'k': case 'M': case 'm': case 'S': case 's':
		case 'W': case 'w': case 'B': case 'b': case 'D': case 'd': case 'H': case 'h': case 'V': case 'v':
			return true;
		default:
			return false;
	}
}

unsigned HostThreads()
{
	static unsigned n = []() {
		if (const char* e = getenv("GA_HOST_THREADS"))
		{
			int v = atoi(e);
			if (v > 0) return (unsigned)v;
		}
		unsigned h = std::thread::hardware_concurrency();
		return h ? h : 1u;
	}();
	return n;
}

// ---- worker pool -------------------------------------------------------------------------------------------------
// The host passes of a batch are a handful of short parallel loops; starting threads for each costs as much as the loop.
// One process-wide pool (never torn down: workers sleep on a condition variable) runs them.  A second caller arriving
// while the pool is busy (one host thread per GPU) falls back to threads of its own.
namespace
{
struct WorkerPool
{
	std::mutex jobMutex;              // one job at a time
	std::mutex m;
	std::condition_variable cvStart, cvDone;
	uint64_t generation = 0;
	const std::function<void()>* job = nullptr;
	unsigned wanted = 0, taken = 0, pending = 0;
	std::vector<std::thread> workers;

	void workerMain()
	{
		uint64_t seen = 0;
		std::unique_lock<std::mutex> lock(m);
		while (true)
		{
			cvStart.wait(lock, [&]() { return generation != seen; });
			seen = generation;
			if (taken >= wanted) continue;
			taken++;
			const std::function<void()>* j = job;
			lock.unlock();
			(*j)();
			lock.lock();
			if (--pending == 0) cvDone.notify_all();
		}
	}
	// runs job on `extra` pool workers and on the caller
	void run(unsigned extra, const std::function<void()>& j)
	{
		{
			std::lock_guard<std::mutex> lock(m);
			while (workers.size() < extra)
			{
				workers.emplace_back([this]() { workerMain(); });
				workers.back().detach();
			}
			job = &j;
			wanted = extra;
			taken = 0;
			pending = extra;
			generation++;
		}
		cvStart.notify_all();
		j();
		std::unique_lock<std::mutex> lock(m);
		cvDone.wait(lock, [&]() { return pending == 0; });
		job = nullptr;
	}
};
WorkerPool& workerPool()
{
	static WorkerPool* pool = new WorkerPool();   // leaked on purpose: detached workers outlive static destruction
	return *pool;
}
}

void ParallelFor(size_t n, const std::function<void(size_t)>& f)
{
	unsigned threads = (unsigned)std::min<size_t>(HostThreads(), n);
	if (threads <= 1)
	{
		for (size_t i = 0; i < n; i++) f(i);
		return;
	}
	std::atomic<size_t> next(0);
	const size_t chunk = std::max<size_t>(1, n / (threads * 8));
	std::exception_ptr error;
	std::atomic<bool> failed(false);
	std::function<void()> work = [&]() {
		try
		{
			while (!failed.load())
			{
				size_t begin = next.fetch_add(chunk);
				if (begin >= n) break;
				size_t end = std::min(n, begin + chunk);
				for (size_t i = begin; i < end; i++) f(i);
			}
		}
		catch (...)
		{
			if (!failed.exchange(true)) error = std::current_exception();
		}
	};
	WorkerPool& wp = workerPool();
	std::unique_lock<std::mutex> jobLock(wp.jobMutex, std::try_to_lock);
	if (jobLock.owns_lock())
	{
		wp.run(threads - 1, work);
	}
	else
	{
		std::vector<std::thread> own;
		for (unsigned t = 0; t + 1 < threads; t++) own.emplace_back(work);
		work();
		for (auto& t : own) t.join();
	}
	if (error) std::rethrow_exception(error);
}

// ---- block cache ---------------------------------------------------------------------------------------------------
// Result buffers are 100 MB-class and live for one batch.  glibc hands such blocks straight to mmap / munmap, so every
// batch would fault all of its pages in again (tens of thousands of faults); freed blocks are kept for the next batch.
namespace
{
struct BlockCache
{
	std::mutex m;
	std::vector<std::pair<void*, size_t>> blocks;   // oldest first
	size_t cached = 0;
	size_t limit;
	BlockCache()
	{
		limit = (size_t)1024 << 20;
		if (const char* e = getenv("GA_HOST_CACHE_MB")) limit = (size_t)std::max(0ll, atoll(e)) << 20;
	}
};
BlockCache& blockCache()
{
	static BlockCache* c = new BlockCache();
	return *c;
}
const size_t kBigBlock = (size_t)1 << 20;
}

void* BigAlloc(size_t bytes, size_t& capOut)
{
	if (bytes < kBigBlock)
	{
		capOut = bytes;
		void* p = malloc(bytes ? bytes : 1);
		if (!p) throw std::bad_alloc();
		return p;
	}
	BlockCache& c = blockCache();
	{
		std::lock_guard<std::mutex> lock(c.m);
		size_t best = c.blocks.size();
		for (size_t i = 0; i < c.blocks.size(); i++)
		{
			const size_t cap = c.blocks[i].second;
			if (cap >= bytes && cap <= 2 * bytes + (8u << 20) && (best == c.blocks.size() || cap < c.blocks[best].second)) best = i;
		}
		if (best != c.blocks.size())
		{
			void* p = c.blocks[best].first;
			capOut = c.blocks[best].second;
			c.cached -= capOut;
			c.blocks.erase(c.blocks.begin() + best);
			return p;
		}
	}
	const size_t round = (size_t)4 << 20;
	capOut = (bytes + bytes / 16 + round - 1) / round * round;
	void* p = malloc(capOut);
	if (!p) throw std::bad_alloc();
	return p;
}

void BigFree(void* p, size_t cap)
{
	if (!p) return;
	if (cap < kBigBlock) { free(p); return; }
	BlockCache& c = blockCache();
	std::vector<void*> drop;
	{
		std::lock_guard<std::mutex> lock(c.m);
		c.blocks.emplace_back(p, cap);
		c.cached += cap;
		while (c.cached > c.limit && !c.blocks.empty())
		{
			drop.push_back(c.blocks.front().first);
			c.cached -= c.blocks.front().second;
			c.blocks.erase(c.blocks.begin());
		}
	}
	for (void* d : drop) free(d);
}

static char complementOf(char c)
{
	// one character of ReverseComplement above
	switch (c)
	{
		case 'A': case 'a': return 'T';
		case 'C': case 'c': return 'G';
		case 'T': case 't': return 'A';
		case 'G': case 'g': return 'C';
		case 'N': case 'n': return 'N';
		case 'U': case 'u': return 'A';
		case 'R': case 'r': return 'Y';
		case 'Y': case 'y': return 'R';
		case 'K': case 'k': return 'M';
		case 'M': case 'm': return 'K';
		case 'S': case 's': return 'S';
		case 'W': case 'w': return 'W';
		case 'B': case 'b': return 'V';
		case 'V': case 'v': return 'B';
		case 'D': case 'd': return 'H';
		case 'H': case 'h': return 'D';
		default: return 'N';
	}
}

BatchPlan::BatchPlan(const AlignmentGraph& graph, const std::vector<ReadInput>& reads, const std::function<uint8_t*(size_t)>& allocParts,
	const std::function<void(size_t, size_t)>& partsReady, bool inPlace)
{
	const size_t n = reads.size();
	badChar.assign(n, 0);
	static const bool deviceMappings = getenv("GA_NO_DEVICE_MAPPINGS") == nullptr;   // env: A/B measurements with the host writing every mapping
	// the reads' bytes go to the device as they are, back to back; a stream names its source range in them and the
	// device reads the (reverse-complemented, padded) part from there.  Which reads hold a character the reference
	// aborts on is found out on the device as well (badChar is filled in when the results come back).
	readOff.resize(n + 1);
	size_t top = 0;
	// in place: the caller's buffer is page-locked and the reads lie in it back to back (what ga_batch's offsets describe) - it
	// is uploaded from where it is, nothing is copied on the host
	if (inPlace && n > 0)
	{
		for (size_t ri = 0; ri + 1 < n && inPlace; ri++) inPlace = reads[ri + 1].seq == reads[ri].seq + reads[ri].seqLen;
	}
	else inPlace = false;
	if (inPlace)
	{
		for (size_t ri = 0; ri < n; ri++) readOff[ri] = (size_t)(reads[ri].seq - reads[0].seq);
		top = readOff[n - 1] + reads[n - 1].seqLen;
	}
	else for (size_t ri = 0; ri < n; ri++) { readOff[ri] = top; top += reads[ri].seqLen; }
	readOff[n] = top;
	partsBytes = top;
	firstSeedOfRead.reserve(n + 1);
	const size_t overlap = (size_t)graph.DBGOverlap;
	for (size_t ri = 0; ri < n; ri++)
	{
		firstSeedOfRead.push_back((uint32_t)seeds.size());
		const size_t len = reads[ri].seqLen;
		for (size_t si = 0; si < reads[ri].nSeeds; si++)
		{
			const SeedHit& hit = reads[ri].seeds[si];
			SeedPlan sp;
			sp.read = (uint32_t)ri;
			sp.seed = (uint32_t)si;
			sp.fwStream = -1;
			sp.bwStream = -1;
			sp.invalid = false;
			int nodeId = std::get<0>(hit);
			size_t pos = std::get<1>(hit);
			bool backwards = std::get<2>(hit);
			if (!graph.HasNode(nodeId * 2) || !graph.HasNode(nodeId * 2 + 1) || pos >= len || pos + overlap > len)
			{
				sp.invalid = true;
				seeds.push_back(sp);
				continue;
			}
			// getSplitAlignment, GraphAligner.h:2969-3024
			size_t forwardNode = graph.Lookup(backwards ? nodeId * 2 + 1 : nodeId * 2);
			size_t backwardNode = graph.Lookup(backwards ? nodeId * 2 : nodeId * 2 + 1);
			if (pos > 0)
			{
				// backward part = reverse complement of read[0 .. pos + overlap), 'N'-padded to a multiple of 64
				size_t partLen = pos + overlap;
				ga_stream_in in;
				in.startNode = (uint32_t)backwardNode;
				in.seqOff = readOff[ri] + partLen - 1;
				in.partLen = (uint32_t)((partLen + 63) / 64 * 64);
				in.trimRows = (uint32_t)pos;                       // GraphAligner.h:3086-3089
				in.srcInfo = (uint32_t)partLen | GA_SRC_BACKWARD;
				sp.bwStream = (int64_t)streams.size();
				streams.push_back(in);
			}
			if (pos < len - 1)
			{
				size_t partLen = len - pos;
				ga_stream_in in;
				in.startNode = (uint32_t)forwardNode;
				in.seqOff = readOff[ri] + pos;
				in.partLen = (uint32_t)((partLen + 63) / 64 * 64);
				in.trimRows = (uint32_t)(len - pos - overlap);     // GraphAligner.h:3063-3066
				in.srcInfo = (uint32_t)partLen;
				// the read's only stream (one seed, at read position 0): its mappings can be written on the device
				if (deviceMappings && reads[ri].nSeeds == 1 && pos == 0 && partLen < 0x40000000u) in.srcInfo |= GA_SRC_SOLO;
				sp.fwStream = (int64_t)streams.size();
				streams.push_back(in);
			}
			seeds.push_back(sp);
		}
	}
	firstSeedOfRead.push_back((uint32_t)seeds.size());
	if (inPlace)
	{
		parts = (uint8_t*)const_cast<char*>(reads[0].seq);
		if (allocParts) allocParts(top + 64);   // the device twin only: the allocator knows that the plan is in place
		if (partsReady)
		{
			// a few ranges, so that the first kernels' input is on its way while the rest is queued
			const size_t pieces = std::min<size_t>(4, std::max<size_t>(1, top >> 22));
			for (size_t k = 0; k < pieces; k++) partsReady(top * k / pieces, top * (k + 1) / pieces - top * k / pieces);
		}
		return;
	}
	if (allocParts) parts = allocParts(top + 64);
	else
	{
		ownedParts.resize(top + 64);
		parts = ownedParts.data();
	}
	// the copy into the (pinned) staging buffer: a few groups of consecutive reads, each handed over when it is complete
	const size_t groups = partsReady ? std::min<size_t>(4, std::max<size_t>(1, n / 64)) : 1;
	size_t groupBegin = 0;
	for (size_t gi = 0; gi < groups; gi++)
	{
		size_t groupEnd = groupBegin;
		const size_t targetBytes = top * (gi + 1) / groups;
		while (groupEnd < n && (gi + 1 == groups || readOff[groupEnd + 1] <= targetBytes)) groupEnd++;
		// pieces of at most 1 MiB so that a few long reads still spread over the workers
		struct Piece { size_t off, bytes; const char* src; };
		std::vector<Piece> pieces;
		for (size_t ri = groupBegin; ri < groupEnd; ri++)
		{
			for (size_t o = 0; o < reads[ri].seqLen; o += (size_t)1 << 20) pieces.push_back(Piece { readOff[ri] + o, std::min<size_t>((size_t)1 << 20, reads[ri].seqLen - o), reads[ri].seq + o });
		}
		ParallelFor(pieces.size(), [&](size_t k) { memcpy(parts + pieces[k].off, pieces[k].src, pieces[k].bytes); });
		if (partsReady && groupEnd > groupBegin && readOff[groupEnd] > readOff[groupBegin]) partsReady(readOff[groupBegin], readOff[groupEnd] - readOff[groupBegin]);
		groupBegin = groupEnd;
	}
}

namespace
{

bool charMatch(char readChar, char graphChar)
{
	// GraphAligner.h:2039-2110
	switch (readChar)
	{
		case 'A': case 'a': return graphChar == 'A';
		case 'T': case 't': return graphChar == 'T';
		case 'C': case 'c': return graphChar == 'C';
		case 'G': case 'g': return graphChar == 'G';
		case 'N': case 'n': return true;
		case 'R': case 'r': return graphChar == 'A' || graphChar == 'G';
		case 'Y': case 'y': return graphChar == 'C' || graphChar == 'T';
		case 'K': case 'k': return graphChar == 'G' || graphChar == 'T';
		case 'M': case 'm': return graphChar == 'C' || graphChar == 'A';
		case 'S': case 's': return graphChar == 'C' || graphChar == 'G';
		case 'W': case 'w': return graphChar == 'A' || graphChar == 'T';
		case 'B': case 'b': return graphChar == 'C' || graphChar == 'G' || graphChar == 'T';
		case 'D': case 'd': return graphChar == 'A' || graphChar == 'G' || graphChar == 'T';
		case 'H': case 'h': return graphChar == 'A' || graphChar == 'C' || graphChar == 'T';
		case 'V': case 'v': return graphChar == 'A' || graphChar == 'C' || graphChar == 'G';
		default: return false;
	}
}


// every position of one stream's trimmed trace in forward order (only needed for TraceItems)
void decodePositions(const AlignmentGraph& graph, const ga_stream_in& in, const ga_stream_out& out, const uint32_t* arena, std::vector<MatrixPos>& t)
{
	t.clear();
	if (out.status != GA_OK || out.nSlices <= 0) return;
	const uint32_t* moves = arena + out.traceOff;
	const uint32_t* path = moves + (out.nMoves + 15) / 16;
	t.reserve(out.nPositions);
	MatrixPos cur;
	cur.node = out.endNode;
	cur.off = out.endOff;
	cur.j = (size_t)out.nSlices * 64 - 1;
	if (cur.j < in.trimRows) t.push_back(cur);
	uint32_t nextPath = 0;
	for (uint32_t m = 0; m < out.nMoves; m++)
	{
		uint32_t move = (moves[m >> 4] >> ((m & 15) * 2)) & 3u;
		if (move == GA_MOVE_END) break;   // the step to row -1 is popped again by the reference (GraphAligner.h:949-951)
		if (move != GA_MOVE_V)
		{
			if (cur.off == 0)
			{
				cur.node = path[nextPath++];
				cur.off = (uint32_t)graph.NodeLength(cur.node) - 1;
			}
			else
			{
				cur.off--;
			}
		}
		if (move != GA_MOVE_H) cur.j--;
		if (cur.j < in.trimRows) t.push_back(cur);
	}
	std::reverse(t.begin(), t.end());
}

AlignmentResult emptyAlignment()
{
	AlignmentResult r;
	r.alignment.score = std::numeric_limits<int32_t>::max();
	r.alignmentFailed = true;
	return r;
}

// getTraceInfoInner, GraphAligner.h:718-780
void traceInfoInner(const AlignmentGraph& graph, const ReadInput& read, const std::vector<MatrixPos>& trace, std::vector<AlignmentResult::TraceItem>& result)
{
	for (size_t i = 1; i < trace.size(); i++)
	{
		const MatrixPos& np = trace[i];
		const MatrixPos& op = trace[i - 1];
		bool sameColumn = np.node == op.node && np.off == op.off;
		bool diagonal = np.j != op.j;
		if (sameColumn)
		{
			// a one-bp node with a self loop may be re-entered diagonally
			if (!(np.j == op.j + 1 && graph.NodeLength(np.node) == 1 && graph.HasOutNeighbor(np.node, np.node))) diagonal = false;
		}
		AlignmentResult::TraceItem item;
		item.nodeID = graph.NodeID(np.node) / 2;
		item.reverse = graph.NodeID(np.node) % 2 == 1;
		item.offset = np.off;
		item.readpos = np.j;
		item.graphChar = graph.NodeSequences(graph.NodeStart(np.node) + np.off);
		item.readChar = np.j < read.seqLen ? read.seq[np.j] : '\0';
		if (np.j == op.j) item.type = AlignmentResult::DELETION;
		else if (sameColumn && !diagonal) item.type = AlignmentResult::INSERTION;
		else item.type = charMatch(item.readChar, item.graphChar) ? AlignmentResult::MATCH : AlignmentResult::MISMATCH;
		result.push_back(item);
	}
}

}

namespace
{

// The same-node runs of one stream's trace in the orientation the read is reported in, straight from the device's record
// (backward order, GA_RUN_WORDS words each), without a copy:
//   forward part   run i = record n-1-i, rows shifted by `shift` (getPiecewiseTracesFromSplit, GraphAligner.h:3090-3093)
//   backward part  reverseTrace (GraphAligner.h:3026-3037): order reversed, positions mapped to the other strand, rows
//                  mirrored around `end` - so run i = record i
struct RunView
{
	const AlignmentGraph* graph = nullptr;
	const uint32_t* rec = nullptr;
	size_t n = 0;
	bool backward = false;
	size_t end = 0, shift = 0;
	uint32_t node(size_t i) const
	{
		if (!backward) return rec[(n - 1 - i) * GA_RUN_WORDS];
		return (uint32_t)graph->GetReverseNode(rec[i * GA_RUN_WORDS]);
	}
	TraceRun get(size_t i) const
	{
		if (!backward)
		{
			const uint32_t* r = rec + (n - 1 - i) * GA_RUN_WORDS;
			return TraceRun { r[0], r[1], r[2], (size_t)r[3] + shift, (size_t)r[4] + shift };
		}
		const uint32_t* r = rec + i * GA_RUN_WORDS;
		const size_t other = graph->GetReverseNode(r[0]);
		const uint32_t len = (uint32_t)graph->NodeLength(other);
		return TraceRun { (uint32_t)other, len - 1 - r[2], len - 1 - r[1], end - (size_t)r[4], end - (size_t)r[3] };
	}
};

RunView viewOf(const AlignmentGraph& graph, const ga_stream_out& out, const uint32_t* arena, bool backward, size_t end, size_t shift)
{
	RunView v;
	v.graph = &graph;
	v.backward = backward;
	v.end = end;
	v.shift = shift;
	if (out.status != GA_OK || out.nSlices <= 0) return v;
	v.rec = arena + out.traceOff + (out.nMoves + 15) / 16 + out.nPathNodes;
	v.n = out.nRuns;
	return v;
}

// traceToAlignment, GraphAligner.h:782-847, from runs: one Mapping per run with exactly one Edit; the runs on the dummy
// nodes at either end are dropped.  Returns false for a failed (empty) direction, else the range [k, last] of runs kept.
bool mappingRange(const AlignmentGraph& graph, const RunView& v, size_t& k, size_t& last)
{
	if (v.n == 0) return false;
	k = 0;
	while (v.node(k) == graph.DummyNodeStart())
	{
		k++;
		if (k == v.n) return false;
	}
	if (v.node(k) == graph.DummyNodeEnd()) return false;
	last = k;
	while (last + 1 < v.n && v.node(last + 1) != graph.DummyNodeEnd()) last++;
	return true;
}

// mapping i (k <= i <= last) of a direction: non-final mappings get from_length = end - start + 1, the final one
// end - start; only the first mapping has an offset
FlatMapping mappingOf(const AlignmentGraph& graph, const ReadInput& read, const RunView& v, size_t k, size_t last, size_t i, size_t& beforeJ)
{
	const TraceRun r = v.get(i);
	FlatMapping m;
	m.rank = (int64_t)(i - k);
	m.node_id = graph.NodeID(r.node);
	m.is_reverse = graph.Reverse(r.node);
	m.offset = i == k ? r.firstOff : 0;
	m.from_length = (int32_t)(r.lastOff - r.firstOff) + (i == last ? 0 : 1);
	m.to_length = (int32_t)(r.lastJ - beforeJ);
	m.read_start = r.firstJ;
	if (r.firstJ > read.seqLen) throw std::out_of_range("basic_string::substr");   // what sequence.substr would do
	beforeJ = r.lastJ;
	return m;
}

}

ReadAssembly AssembleRead(const AlignmentGraph& graph, const ReadInput& read, const BatchPlan& plan, uint32_t readIndex,
	const ga_stream_out* outs, const uint32_t* arena, bool materialize)
{
	ReadAssembly as;
	uint32_t first = plan.firstSeedOfRead[readIndex], last = plan.firstSeedOfRead[readIndex + 1];
	for (uint32_t k = first; k < last; k++)
	{
		const BatchPlan::SeedPlan& sp = plan.seeds[k];
		if (sp.fwStream >= 0) as.wordColumns += outs[sp.fwStream].wordColumns;
		if (sp.bwStream >= 0) as.wordColumns += outs[sp.bwStream].wordColumns;
	}
	// reference: abort on a character outside its IUPAC switch, before any seed is looked at
	if (plan.badChar[readIndex] && last > first)
	{
		as.flags |= FLAG_BAD_CHAR;
		return as;
	}
	// A read whose single stream carries device-written mapping records (one valid seed at read position 0): everything the seed
	// loop below would derive from its runs is already there - no backward part, rows not shifted, nothing to merge or prune.
	if (last - first == 1 && !plan.seeds[first].invalid && plan.seeds[first].bwStream < 0 && plan.seeds[first].fwStream >= 0)
	{
		const ga_stream_out& o = outs[plan.seeds[first].fwStream];
		if (o.status == GA_OK && o.nSlices > 0 && o.nMapped > 0)
		{
			if (o.cyclicSlices) as.flags |= FLAG_CYCLIC;
			if (o.rampRedos) as.flags |= FLAG_RAMP_REDO | ((o.rampRedos & GA_RAMP_STALE_BIT) ? FLAG_RAMP_STALE : 0u);
			const uint64_t recWord = o.traceOff + (o.nMoves + 15) / 16 + o.nPathNodes;
			as.deviceMapped = true;
			as.deviceMapWord = recWord + GA_MAP_PAD(recWord);
			as.nMappings = o.nMapped;
			as.score = o.score;
			as.fwStream = o.nPositions > 0 ? plan.seeds[first].fwStream : -1;
			as.mapFwStream = plan.seeds[first].fwStream;
			as.splitIndex = std::get<1>(read.seeds[plan.seeds[first].seed]);   // 0
			as.fwShifted = false;
			as.queryPosition = (int32_t)as.splitIndex;
			as.alignmentStart = as.splitIndex;
			as.alignmentEnd = as.splitIndex + (size_t)o.nSlices * 64;
			as.estimated = (size_t)o.nSlices * 64;
			as.nTraceItems = o.nPositions > 0 ? o.nPositions - 1 : 0;
			as.failed = false;
			if (materialize)
			{
				as.mappings.resize(as.nMappings);
				EmitMappings(graph, read, as, outs, arena, [&](size_t i, const FlatMapping& m) { as.mappings[i] = m; });
			}
			return as;
		}
	}
	std::vector<std::tuple<size_t, size_t, size_t>> tried;
	bool hasAlignment = false;
	RunView bestFw, bestBw;
	int32_t bestFwScore = 0, bestBwScore = 0;
	size_t bestEstimated = 0, bestSeedPos = 0, bestFwN = 0, bestBwN = 0;
	for (uint32_t k = first; k < last; k++)
	{
		const BatchPlan::SeedPlan& sp = plan.seeds[k];
		const SeedHit& hit = read.seeds[sp.seed];
		if (sp.invalid)
		{
			// reference: nodeLookup.at / substr throw std::out_of_range (GraphAligner.h:423)
			as.flags |= FLAG_BAD_SEED;
			return as;
		}
		size_t nodeIndex = graph.Lookup(std::get<0>(hit) * 2);
		size_t pos = std::get<1>(hit);
		bool already = false;
		for (auto& t : tried)
		{
			if (std::get<0>(t) <= pos && std::get<1>(t) >= pos && std::get<2>(t) == nodeIndex) { already = true; break; }
		}
		if (already) continue;   // "seed i already aligned", GraphAligner.h:425-429
		// getPiecewiseTracesFromSplit, GraphAligner.h:3039-3098
		const size_t splitIndex = pos;
		bool streamError = false, shifted = false;
		size_t fwSlices = 0, bwSlices = 0, fwN = 0, bwN = 0;
		int32_t fwScore = 0, bwScore = 0;
		RunView fw, bw;
		if (sp.bwStream >= 0)
		{
			const ga_stream_out& o = outs[sp.bwStream];
			if (o.status != GA_OK && o.status != GA_EMPTY) streamError = true;
			if (o.cyclicSlices) as.flags |= FLAG_CYCLIC;
			if (o.rampRedos) as.flags |= FLAG_RAMP_REDO | ((o.rampRedos & GA_RAMP_STALE_BIT) ? FLAG_RAMP_STALE : 0u);
			if (o.status == GA_OK && o.nSlices > 0)
			{
				bwSlices = (size_t)o.nSlices;
				bwScore = o.score;
				bwN = o.nPositions;
				bw = viewOf(graph, o, arena, true, splitIndex - 1, 0);
				// the forward rows are shifted only inside this branch in the reference (GraphAligner.h:3090-3093)
				shifted = true;
			}
		}
		if (sp.fwStream >= 0)
		{
			const ga_stream_out& o = outs[sp.fwStream];
			if (o.status != GA_OK && o.status != GA_EMPTY) streamError = true;
			if (o.cyclicSlices) as.flags |= FLAG_CYCLIC;
			if (o.rampRedos) as.flags |= FLAG_RAMP_REDO | ((o.rampRedos & GA_RAMP_STALE_BIT) ? FLAG_RAMP_STALE : 0u);
			if (o.status == GA_OK && o.nSlices > 0)
			{
				fwSlices = (size_t)o.nSlices;
				fwScore = o.score;
				fwN = o.nPositions;
				fw = viewOf(graph, o, arena, false, 0, shifted ? splitIndex : 0);
			}
		}
		if (streamError)
		{
			as.flags |= FLAG_STREAM_ERROR;
			return as;
		}
		size_t estimated = (fwSlices + bwSlices) * 64;
		// addAlignmentNodes, GraphAligner.h:594-634
		if (k + 1 < last)   // only a later seed of this read ever looks at them
		{
			tried.reserve(tried.size() + fw.n + bw.n);
			for (size_t i = 0; i < fw.n; i++) { const TraceRun r = fw.get(i); tried.emplace_back(r.firstJ, r.lastJ, (size_t)r.node); }
			for (size_t i = 0; i < bw.n; i++) { const TraceRun r = bw.get(i); tried.emplace_back(r.firstJ, r.lastJ, (size_t)r.node); }
		}
		if (!hasAlignment || estimated > bestEstimated)
		{
			bestFw = fw;
			bestBw = bw;
			bestFwScore = fwScore;
			bestBwScore = bwScore;
			bestFwN = fwN;
			bestBwN = bwN;
			bestEstimated = estimated;
			bestSeedPos = pos;
			hasAlignment = true;
			as.fwStream = fwN > 0 ? sp.fwStream : -1;
			as.bwStream = bwN > 0 ? sp.bwStream : -1;
			as.mapFwStream = fw.n > 0 ? sp.fwStream : -1;
			as.mapBwStream = bw.n > 0 ? sp.bwStream : -1;
			as.splitIndex = splitIndex;
			as.fwShifted = shifted;
		}
	}
	if (!hasAlignment) return as;
	as.estimated = bestEstimated;
	size_t fk = 0, fl = 0, bk = 0, bl = 0;
	const bool fwOk = mappingRange(graph, bestFw, fk, fl);
	const bool bwOk = mappingRange(graph, bestBw, bk, bl);
	if (!fwOk && !bwOk) return as;
	// mergeAlignments, GraphAligner.h:648-688: bw first, then fw without its first mapping when both meet on the same node
	// (otherwise the reference checks the edge and, failing that, only logs "Piecewise alignments can't be merged!")
	size_t start = 0;
	if (fwOk && bwOk)
	{
		as.score = bestBwScore + bestFwScore;
		const uint32_t a = bestBw.node(bl), b2 = bestFw.node(fk);
		if (graph.NodeID(a) == graph.NodeID(b2) && graph.Reverse(a) == graph.Reverse(b2)) start = 1;
	}
	else as.score = bwOk ? bestBwScore : bestFwScore;
	as.mapBwFirst = bk;
	as.mapBwCount = bwOk ? bl - bk + 1 : 0;
	as.mapFwFirst = fk;
	as.mapFwCount = fwOk ? fl - fk + 1 : 0;
	as.mapFwSkip = start;
	as.nMappings = as.mapBwCount + as.mapFwCount - start;
	size_t lastAligned = bestBw.n > 0 ? bestBw.get(0).firstJ : bestSeedPos;
	as.queryPosition = (int32_t)lastAligned;
	as.alignmentStart = lastAligned;
	as.alignmentEnd = lastAligned + bestEstimated;
	as.nTraceItems = (bestBwN > 0 ? bestBwN - 1 : 0) + ((bestBwN > 0 && bestFwN > 0) ? 1 : 0) + (bestFwN > 0 ? bestFwN - 1 : 0);
	as.failed = false;
	if (materialize)
	{
		as.mappings.resize(as.nMappings);
		EmitMappings(graph, read, as, outs, arena, [&](size_t i, const FlatMapping& m) { as.mappings[i] = m; });
	}
	return as;
}

// ---- seeds in two rounds (ga_align_batch): what a later seed of a read needs to know about an earlier seed's alignment ----
bool SeedIsValid(const AlignmentGraph& graph, const ReadInput& read, const SeedHit& hit)
{
	// the checks of BatchPlan (reference: nodeLookup.at / substr throw std::out_of_range, GraphAligner.h:423,2977-2998)
	const int nodeId = std::get<0>(hit);
	const size_t pos = std::get<1>(hit);
	return graph.HasNode(nodeId * 2) && graph.HasNode(nodeId * 2 + 1) && pos < read.seqLen && pos + (size_t)graph.DBGOverlap <= read.seqLen;
}

void CollectTried(const AlignmentGraph& graph, const ga_stream_out* outs, const uint32_t* arena, int64_t fwStream, int64_t bwStream, size_t splitIndex, bool fwShifted,
	std::vector<std::tuple<size_t, size_t, size_t>>& tried)
{
	// addAlignmentNodes, GraphAligner.h:594-634: (first row, last row, node) of every run of the seed's forward and backward trace
	if (fwStream >= 0)
	{
		const ga_stream_out& o = outs[fwStream];
		if (o.status == GA_OK && o.nSlices > 0 && o.nMapped > 0)
		{
			// the device wrote mapping records instead of runs (rows not shifted): the rows follow from read_start / to_length
			const uint64_t recWord = o.traceOff + (o.nMoves + 15) / 16 + o.nPathNodes;
			const GaDeviceMapping* rec = (const GaDeviceMapping*)(arena + recWord + GA_MAP_PAD(recWord));
			size_t lastJ = rec[0].read_start;
			for (uint32_t i = 0; i < o.nMapped; i++)
			{
				lastJ += (size_t)rec[i].to_length;
				tried.emplace_back((size_t)rec[i].read_start, lastJ, graph.Lookup((int)rec[i].node_id));
			}
		}
		else
		{
			const RunView fw = viewOf(graph, o, arena, false, 0, fwShifted ? splitIndex : 0);
			for (size_t i = 0; i < fw.n; i++) { const TraceRun r = fw.get(i); tried.emplace_back(r.firstJ, r.lastJ, (size_t)r.node); }
		}
	}
	if (bwStream >= 0)
	{
		const RunView bw = viewOf(graph, outs[bwStream], arena, true, splitIndex - 1, 0);
		for (size_t i = 0; i < bw.n; i++) { const TraceRun r = bw.get(i); tried.emplace_back(r.firstJ, r.lastJ, (size_t)r.node); }
	}
}

bool SeedCovered(const AlignmentGraph& graph, const std::vector<std::tuple<size_t, size_t, size_t>>& tried, const SeedHit& hit)
{
	// "seed i already aligned", GraphAligner.h:425-429
	const size_t nodeIndex = graph.Lookup(std::get<0>(hit) * 2);
	const size_t pos = std::get<1>(hit);
	for (const auto& t : tried)
	{
		if (std::get<0>(t) <= pos && std::get<1>(t) >= pos && std::get<2>(t) == nodeIndex) return true;
	}
	return false;
}

namespace
{
template <typename Sink>
void emitMappingsTo(const AlignmentGraph& graph, const ReadInput& read, const ReadAssembly& as, const ga_stream_out* outs, const uint32_t* arena, const Sink& sink)
{
	if (as.failed) return;
	if (as.deviceMapped)
	{
		const GaDeviceMapping* rec = (const GaDeviceMapping*)(arena + as.deviceMapWord);
		for (size_t i = 0; i < as.nMappings; i++)
		{
			FlatMapping m;
			m.node_id = rec[i].node_id;
			m.is_reverse = rec[i].is_reverse != 0;
			m.offset = rec[i].offset;
			m.rank = rec[i].rank;
			m.from_length = rec[i].from_length;
			m.to_length = rec[i].to_length;
			m.read_start = rec[i].read_start;
			if (m.read_start > read.seqLen) throw std::out_of_range("basic_string::substr");   // what sequence.substr would do
			sink(i, m);
		}
		return;
	}
	size_t idx = 0;
	if (as.mapBwCount)
	{
		const RunView v = viewOf(graph, outs[as.mapBwStream], arena, true, as.splitIndex - 1, 0);
		const size_t k = as.mapBwFirst, last = k + as.mapBwCount - 1;
		size_t beforeJ = v.get(k).firstJ;
		for (size_t i = k; i <= last; i++) sink(idx++, mappingOf(graph, read, v, k, last, i, beforeJ));
	}
	if (as.mapFwCount)
	{
		const RunView v = viewOf(graph, outs[as.mapFwStream], arena, false, 0, as.fwShifted ? as.splitIndex : 0);
		const size_t k = as.mapFwFirst, last = k + as.mapFwCount - 1;
		size_t beforeJ = v.get(k).firstJ;
		for (size_t i = k; i <= last; i++)
		{
			const FlatMapping m = mappingOf(graph, read, v, k, last, i, beforeJ);
			if (i - k >= as.mapFwSkip) sink(idx++, m);
		}
	}
}
}

void EmitMappings(const AlignmentGraph& graph, const ReadInput& read, const ReadAssembly& as, const ga_stream_out* outs, const uint32_t* arena,
	const std::function<void(size_t, const FlatMapping&)>& sink)
{
	emitMappingsTo(graph, read, as, outs, arena, sink);
}

void WriteMappings(const AlignmentGraph& graph, const ReadInput& read, const ReadAssembly& as, const ga_stream_out* outs, const uint32_t* arena, ::ga_mapping* dst)
{
	if (as.failed) return;
	if (as.deviceMapped)
	{
		memcpy(dst, arena + as.deviceMapWord, as.nMappings * sizeof(::ga_mapping));   // only when the records have to move (merged results)
		return;
	}
	if (!as.mapBwCount && as.mapFwCount && !as.mapFwSkip)
	{
		// the common case (a seed at read position 0: forward part only) as one flat loop over the device's run records, last run
		// first (mappingOf / RunView::get with backward = false written out)
		const ga_stream_out& out = outs[as.mapFwStream];
		const size_t n = out.nRuns, shift = as.fwShifted ? as.splitIndex : 0;
		const uint32_t* rec = arena + out.traceOff + (out.nMoves + 15) / 16 + out.nPathNodes;
		const size_t k = as.mapFwFirst, count = as.mapFwCount;
		const uint32_t* r = rec + (n - 1 - k) * GA_RUN_WORDS;
		size_t beforeJ = (size_t)r[3] + shift;
		for (size_t i = 0; i < count; i++, r -= GA_RUN_WORDS)
		{
			const uint32_t node = r[0];
			const size_t firstJ = (size_t)r[3] + shift, lastJ = (size_t)r[4] + shift;
			if (firstJ > read.seqLen) throw std::out_of_range("basic_string::substr");
			::ga_mapping& gm = dst[i];
			gm.node_id = graph.NodeID(node);
			gm.offset = i == 0 ? r[1] : 0;
			gm.rank = (uint32_t)i;
			gm.from_length = (int32_t)(r[2] - r[1]) + (i + 1 == count ? 0 : 1);
			gm.to_length = (int32_t)(lastJ - beforeJ);
			gm.read_start = (uint32_t)firstJ;
			gm.is_reverse = graph.Reverse(node) ? 1u : 0u;
			beforeJ = lastJ;
		}
		return;
	}
	emitMappingsTo(graph, read, as, outs, arena, [dst](size_t k, const FlatMapping& m) {
		::ga_mapping& gm = dst[k];
		gm.node_id = m.node_id;
		gm.offset = (uint32_t)m.offset;
		gm.rank = (uint32_t)m.rank;
		gm.from_length = m.from_length;
		gm.to_length = m.to_length;
		gm.read_start = (uint32_t)m.read_start;
		gm.is_reverse = m.is_reverse ? 1u : 0u;
	});
}

AlignmentResult ToAlignmentResult(const ReadInput& read, const ReadAssembly& as, bool keepSequences)
{
	AlignmentResult r = emptyAlignment();
	r.flags = as.flags;
	r.wordColumns = as.wordColumns;
	if (as.failed) return r;
	r.alignmentFailed = false;
	r.alignment.score = as.score;
	r.alignment.query_position = as.queryPosition;
	r.alignmentStart = as.alignmentStart;
	r.alignmentEnd = as.alignmentEnd;
	r.alignment.name.assign(read.name, read.nameLen);
	if (keepSequences) r.alignment.sequence.assign(read.seq, read.seqLen);
	r.alignment.path.mapping.resize(as.mappings.size());
	for (size_t i = 0; i < as.mappings.size(); i++)
	{
		const FlatMapping& f = as.mappings[i];
		vg::Mapping& m = r.alignment.path.mapping[i];
		m.rank = f.rank;
		m.position.node_id = f.node_id;
		m.position.offset = f.offset;
		m.position.is_reverse = f.is_reverse;
		vg::Edit e;
		e.from_length = f.from_length;
		e.to_length = f.to_length;
		e.read_start = f.read_start;
		if (keepSequences && f.read_start <= read.seqLen) e.sequence.assign(read.seq + f.read_start, std::min<size_t>((size_t)std::max(0, f.to_length), read.seqLen - f.read_start));
		m.edit.push_back(std::move(e));
	}
	return r;
}

void BuildTraceItems(const AlignmentGraph& graph, const ReadInput& read, const ReadAssembly& as, const ga_stream_in* streams,
	const ga_stream_out* outs, const uint32_t* arena, std::vector<AlignmentResult::TraceItem>& items)
{
	// getTraceInfo, GraphAligner.h:690-716
	items.clear();
	if (as.failed) return;
	std::vector<MatrixPos> fw, bw;
	if (as.fwStream >= 0)
	{
		decodePositions(graph, streams[as.fwStream], outs[as.fwStream], arena, fw);
		if (as.fwShifted) for (auto& p : fw) p.j += as.splitIndex;
	}
	if (as.bwStream >= 0)
	{
		decodePositions(graph, streams[as.bwStream], outs[as.bwStream], arena, bw);
		std::reverse(bw.begin(), bw.end());
		for (auto& p : bw)
		{
			size_t other = graph.GetReverseNode(p.node);
			p.off = (uint32_t)(graph.NodeLength(other) - 1 - p.off);
			p.node = (uint32_t)other;
			p.j = (as.splitIndex - 1) - p.j;
		}
	}
	items.reserve(as.nTraceItems);
	if (!bw.empty()) traceInfoInner(graph, read, bw, items);
	if (!bw.empty() && !fw.empty())
	{
		const MatrixPos& p = fw[0];
		AlignmentResult::TraceItem item;
		item.type = AlignmentResult::FORWARDBACKWARDSPLIT;
		item.nodeID = graph.NodeID(p.node) / 2;
		item.reverse = p.node % 2 == 1;   // node INDEX parity, as the reference writes it (GraphAligner.h:704)
		item.offset = p.off;
		item.readpos = p.j;
		item.graphChar = graph.NodeSequences(graph.NodeStart(p.node) + p.off);
		item.readChar = p.j < read.seqLen ? read.seq[p.j] : '\0';
		items.push_back(item);
	}
	if (!fw.empty()) traceInfoInner(graph, read, fw, items);
}

std::vector<AlignmentResult> AlignBatch(DeviceCtx* ctx, const AlignmentGraph& graph, const std::vector<ReadInput>& reads, int initialBandwidth, int rampBandwidth, BatchStats* stats,
	std::mutex* gpuTurn)
{
	if (!graph.Finalized()) throw std::logic_error("AlignBatch: graph not finalized");
	BatchPlan plan(graph, reads);
	RawBuffer<ga_stream_out> outs;
	RawBuffer<uint32_t> arena;
	{
		// several contexts of one GPU take turns on the device part; planning and assembly overlap with the other's kernel
		std::unique_lock<std::mutex> turn;
		if (gpuTurn) turn = std::unique_lock<std::mutex>(*gpuTurn);
		ExecuteStreams(ctx, plan.streams, plan.parts, plan.partsBytes, plan.readOff, initialBandwidth, rampBandwidth, outs, arena, plan.badChar, stats);
	}
	std::vector<AlignmentResult> results(reads.size());
	ParallelFor(reads.size(), [&](size_t i) {
		if (reads[i].nSeeds == 0)
		{
			results[i] = emptyAlignment();   // Aligner.cpp:131-138 "has no seed hits"
			return;
		}
		ReadAssembly as = AssembleRead(graph, reads[i], plan, (uint32_t)i, outs.data(), arena.data());
		results[i] = ToAlignmentResult(reads[i], as, true);
		BuildTraceItems(graph, reads[i], as, plan.streams.data(), outs.data(), arena.data(), results[i].trace);
	});
	if (stats)
	{
		stats->streams += plan.streams.size();
		for (size_t i = 0; i < outs.size(); i++) stats->wordColumns += outs.data()[i].wordColumns;
	}
	return results;
}

}
