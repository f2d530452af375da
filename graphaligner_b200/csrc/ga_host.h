// Host side of the hot path: turns reads + seed hits into DP streams for the GPU and the GPU's traces back
// into the reference's AlignmentResult (reference GraphAlignerWrapper.h:10-54, GraphAligner.h:408-491).
#ifndef GA_HOST_H
#define GA_HOST_H
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <tuple>
#include <mutex>
#include <vector>
#include "alignment_graph.h"
#include "ga_types.h"

namespace vg
{
// plain-struct stand-ins for the vg.proto messages the aligner fills in (vg.pb.h:149-960)
struct Position
{
	int64_t node_id = 0;
	int64_t offset = 0;
	bool is_reverse = false;
};
struct Edit
{
	int32_t from_length = 0;
	int32_t to_length = 0;
	std::string sequence;
	size_t read_start = 0;   // not a vg field: sequence == read.substr(read_start, to_length)
};
struct Mapping
{
	Position position;
	std::vector<Edit> edit;
	int64_t rank = 0;
};
struct Path
{
	std::vector<Mapping> mapping;
};
struct Alignment
{
	std::string sequence;
	Path path;
	std::string name;
	int32_t score = 0;
	int32_t query_position = 0;
};
}

class AlignmentResult
{
public:
	enum TraceMatchType
	{
		MATCH = 1,
		MISMATCH = 2,
		INSERTION = 3,
		DELETION = 4,
		FORWARDBACKWARDSPLIT = 5
	};
	struct TraceItem
	{
		int nodeID;
		size_t offset;
		bool reverse;
		size_t readpos;
		TraceMatchType type;
		char graphChar;
		char readChar;
	};
	vg::Alignment alignment;
	bool alignmentFailed = true;
	size_t cellsProcessed = 0;
	size_t elapsedMilliseconds = 0;
	size_t alignmentStart = 0;
	size_t alignmentEnd = 0;
	std::vector<TraceItem> trace;
	// extras (not in the reference): forward-pass word updates of every stream run for this read, and
	// whether any band held a cyclic component / the final slice had a cross-node tie
	uint64_t wordColumns = 0;
	uint32_t flags = 0;
};

struct ga_mapping;   // include/graphaligner_b200.h

namespace ga
{

struct DeviceCtx;   // defined in ga_kernels.cu

typedef std::tuple<int, size_t, bool> SeedHit;   // (bigraph node id, read position, reverse)

// one read of a batch; nothing is copied, the caller's buffers must outlive the batch
struct ReadInput
{
	const char* name;
	size_t nameLen;
	const char* seq;
	size_t seqLen;
	const SeedHit* seeds;
	size_t nSeeds;
};

// 1 MiB+ blocks are recycled through a process-wide cache instead of going back to the OS (ga_host.cpp)
void* BigAlloc(size_t bytes, size_t& capOut);
void BigFree(void* p, size_t cap);

// heap buffer without value-initialisation (a 100 MB std::vector costs ~25 ms just to zero) for trivially copyable T.
// adopt() takes over a block from elsewhere (pinned host memory a device copy landed in) together with its release function.
template <typename T>
class RawBuffer
{
public:
	typedef void (*ReleaseFn)(void* p, size_t capBytes);
	RawBuffer() : p(nullptr), n(0), cap(0), rel(nullptr) {}
	~RawBuffer() { clear(); }
	RawBuffer(const RawBuffer&) = delete;
	RawBuffer& operator=(const RawBuffer&) = delete;
	RawBuffer(RawBuffer&& o) noexcept : p(o.p), n(o.n), cap(o.cap), rel(o.rel) { o.p = nullptr; o.n = 0; o.cap = 0; o.rel = nullptr; }
	RawBuffer& operator=(RawBuffer&& o) noexcept { if (this != &o) { clear(); swap(o); } return *this; }
	void resize(size_t count)
	{
		if (!rel && p && count * sizeof(T) <= cap) { n = count; return; }
		clear();
		if (count) p = (T*)BigAlloc(count * sizeof(T), cap);
		n = count;
	}
	void adopt(T* block, size_t count, size_t capBytes, ReleaseFn release)
	{
		clear();
		p = block; n = count; cap = capBytes; rel = release;
	}
	T* data() { return p; }
	const T* data() const { return p; }
	size_t size() const { return n; }
	void clear()
	{
		if (p) { if (rel) rel(p, cap); else BigFree(p, cap); }
		p = nullptr; n = 0; cap = 0; rel = nullptr;
	}
	void swap(RawBuffer& o) { std::swap(p, o.p); std::swap(n, o.n); std::swap(cap, o.cap); std::swap(rel, o.rel); }
private:
	T* p;
	size_t n, cap;
	ReleaseFn rel;
};

struct MatrixPos
{
	uint32_t node;
	uint32_t off;
	size_t j;
};

// maximal run of consecutive trace positions on the same node (what traceToAlignment turns into one Mapping)
struct TraceRun
{
	uint32_t node;
	uint32_t firstOff, lastOff;
	size_t firstJ, lastJ;
};

struct BatchStats
{
	uint64_t wordColumns = 0;
	uint64_t streams = 0;
	uint64_t retries = 0;
	uint64_t h2dBytes = 0;
	uint64_t d2hBytes = 0;
	uint64_t launches = 0;
	double kernelMs = 0;
	double peqMs = 0, forwardMs = 0, traceMs = 0;   // device time per kernel (events on the context's stream)
};

// Reference reverse complement incl. its IUPAC table and the 'H' fall-through (CommonUtils.cpp:60-136).
std::string ReverseComplement(const std::string& s);
bool ValidReadChar(char c);
unsigned HostThreads();   // worker threads for the host-side passes (GA_HOST_THREADS or hardware_concurrency)

// Plans the streams of a batch (two per seed: backward part, forward part) and stages the reads' bytes for the upload.
class BatchPlan
{
public:
	// allocParts(bytes) may hand out pinned host memory for the padded parts (it must stay valid until the batch has
	// been uploaded); with no allocator the plan owns the storage
	// allocParts: where the parts are built (pinned memory of a device context); partsReady(offset, bytes): called as soon as a
	// contiguous range of them is complete, so that its upload overlaps the building of the rest
	// inPlace: the reads' buffer is page-locked host memory - it becomes `parts` itself (allocParts(0) is still called once, partsReady
	// for ranges of the caller's buffer)
	BatchPlan(const AlignmentGraph& graph, const std::vector<ReadInput>& reads, const std::function<uint8_t*(size_t)>& allocParts = nullptr,
		const std::function<void(size_t, size_t)>& partsReady = nullptr, bool inPlace = false);
	std::vector<ga_stream_in> streams;
	uint8_t* parts = nullptr;       // the reads' bytes, back to back (the streams' parts are ranges of them, ga_stream_in::seqOff / srcInfo)
	size_t partsBytes = 0;
	std::vector<uint64_t> readOff;  // size reads+1: where each read lies in `parts`
	RawBuffer<uint8_t> ownedParts;
	struct SeedPlan
	{
		uint32_t read;
		uint32_t seed;
		int64_t fwStream;   // -1 if the direction does not exist
		int64_t bwStream;
		bool invalid;       // unknown node / position outside the read / bad character
	};
	std::vector<SeedPlan> seeds;
	std::vector<uint32_t> firstSeedOfRead;   // size reads+1
	std::vector<uint8_t> badChar;            // per read: holds a character the reference aborts on (found on the device, valid once the results are back)
};

// What the seeded AlignOneWay decided for one read (GraphAligner.h:408-491): enough to build the
// vg::Alignment now and the TraceItems later, on demand.
// one vg::Mapping with its single Edit, flattened (see runsToAlignment)
struct FlatMapping
{
	int64_t node_id;
	int64_t offset;
	int64_t rank;
	int32_t from_length;
	int32_t to_length;
	uint64_t read_start;
	bool is_reverse;
};

struct ReadAssembly
{
	bool failed = true;
	uint32_t flags = 0;
	uint64_t wordColumns = 0;
	int64_t fwStream = -1, bwStream = -1;   // streams of the chosen seed that contributed a trace
	size_t splitIndex = 0;
	bool fwShifted = false;                 // forward rows were shifted by splitIndex (GraphAligner.h:3090-3093)
	size_t nTraceItems = 0;
	int32_t score = 0x7fffffff;
	int32_t queryPosition = 0;
	size_t alignmentStart = 0, alignmentEnd = 0;
	// the path: mappings [mapBwFirst, +mapBwCount) of the backward stream's runs, then those of the forward stream without
	// its first mapFwSkip (mergeAlignments); EmitMappings writes them out, `mappings` holds them only when asked for
	int64_t mapFwStream = -1, mapBwStream = -1;
	size_t mapBwFirst = 0, mapBwCount = 0, mapFwFirst = 0, mapFwCount = 0, mapFwSkip = 0;
	size_t nMappings = 0;
	size_t estimated = 0;                   // EstimatedCorrectlyAligned of the chosen seed (64 x retained slices), also when the read failed
	// the device wrote the read's mapping records itself (GA_SRC_SOLO stream): they lie in the arena at word deviceMapWord
	// (a multiple of 8), nMappings of them, as GaDeviceMapping = ::ga_mapping
	bool deviceMapped = false;
	uint64_t deviceMapWord = 0;
	std::vector<FlatMapping> mappings;
};

// the reference-shaped AlignmentResult (vg::Alignment with names / sequences / edits) of an assembled read
AlignmentResult ToAlignmentResult(const ReadInput& read, const ReadAssembly& as, bool keepSequences);

// materialize = false: the mappings are only counted (nMappings); EmitMappings produces them later, straight into their
// final place
ReadAssembly AssembleRead(const AlignmentGraph& graph, const ReadInput& read, const BatchPlan& plan, uint32_t readIndex,
	const ga_stream_out* outs, const uint32_t* arena, bool materialize = true);
void EmitMappings(const AlignmentGraph& graph, const ReadInput& read, const ReadAssembly& as, const ga_stream_out* outs, const uint32_t* arena,
	const std::function<void(size_t, const FlatMapping&)>& sink);
// Seeds in two rounds (ga_align_batch): validity of a seed (BatchPlan's checks), the (first row, last row, node) triples of an
// aligned seed's traces (addAlignmentNodes, GraphAligner.h:594-634) and the reference's "seed already aligned" test against them
bool SeedIsValid(const AlignmentGraph& graph, const ReadInput& read, const SeedHit& hit);
void CollectTried(const AlignmentGraph& graph, const ga_stream_out* outs, const uint32_t* arena, int64_t fwStream, int64_t bwStream, size_t splitIndex, bool fwShifted,
	std::vector<std::tuple<size_t, size_t, size_t>>& tried);
bool SeedCovered(const AlignmentGraph& graph, const std::vector<std::tuple<size_t, size_t, size_t>>& tried, const SeedHit& hit);
// the same, as the C ABI's records (include/graphaligner_b200.h), into dst[0 .. nMappings)
void WriteMappings(const AlignmentGraph& graph, const ReadInput& read, const ReadAssembly& as, const ga_stream_out* outs, const uint32_t* arena, ::ga_mapping* dst);

// AlignmentResult::trace of an assembled read (getTraceInfo, GraphAligner.h:690-780), decoded from the device's move record
void BuildTraceItems(const AlignmentGraph& graph, const ReadInput& read, const ReadAssembly& as, const ga_stream_in* streams,
	const ga_stream_out* outs, const uint32_t* arena, std::vector<AlignmentResult::TraceItem>& items);

// Implemented by the CUDA translation unit: runs all streams on the device behind ctx.
void ExecuteStreams(DeviceCtx* ctx, const std::vector<ga_stream_in>& streams, const uint8_t* parts, size_t partsBytes, const std::vector<uint64_t>& readOff, int initialBandwidth, int rampBandwidth,
	RawBuffer<ga_stream_out>& outs, RawBuffer<uint32_t>& arena, std::vector<uint8_t>& badChar, BatchStats* stats);

// C++ batch entry: full AlignmentResults including trace items (the C ABI materialises those lazily instead)
std::vector<AlignmentResult> AlignBatch(DeviceCtx* ctx, const AlignmentGraph& graph, const std::vector<ReadInput>& reads, int initialBandwidth, int rampBandwidth, BatchStats* stats,
	std::mutex* gpuTurn = nullptr);

// runs f(i) for i in [0,n) on HostThreads() workers
void ParallelFor(size_t n, const std::function<void(size_t)>& f);

}

#endif
