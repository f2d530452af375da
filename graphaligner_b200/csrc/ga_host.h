// Host side of the hot path: turns reads + seed hits into DP streams for the GPU and the GPU's traces back
// into the reference's AlignmentResult (reference GraphAlignerWrapper.h:10-54, GraphAligner.h:408-491).
#ifndef GA_HOST_H
#define GA_HOST_H
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
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

namespace ga
{

struct DeviceCtx;   // defined in ga_kernels.cu

typedef std::tuple<int, size_t, bool> SeedHit;   // (bigraph node id, read position, reverse)

struct ReadInput
{
	const std::string* name;
	const std::string* sequence;
	const std::vector<SeedHit>* seeds;
};

struct MatrixPos
{
	uint32_t node;
	uint32_t off;
	size_t j;
};

struct DirectionTrace
{
	int32_t score = 0;
	bool present = false;          // the direction was run and kept >= 1 slice
	size_t nSlices = 0;
	std::vector<MatrixPos> trace;  // ascending rows, like getTraceFromTable's result
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
};

// Reference reverse complement incl. its IUPAC table and the 'H' fall-through (CommonUtils.cpp:60-136).
std::string ReverseComplement(const std::string& s);
bool ValidReadChar(char c);

// Plans the streams of a batch (two per seed: backward part, forward part).
class BatchPlan
{
public:
	BatchPlan(const AlignmentGraph& graph, const std::vector<ReadInput>& reads);
	std::vector<ga_stream_in> streams;
	std::vector<uint8_t> parts;
	struct SeedPlan
	{
		uint32_t read;
		uint32_t seed;
		int64_t fwStream;   // -1 if the direction does not exist
		int64_t bwStream;
		bool invalid;       // unknown node / position outside the read: reference throws out_of_range
	};
	std::vector<SeedPlan> seeds;
	std::vector<uint32_t> firstSeedOfRead;   // size reads+1
};

// Rebuilds the (node, offset, row) trace of one stream from the device's move/path record.
DirectionTrace DecodeStream(const AlignmentGraph& graph, const ga_stream_in& in, const ga_stream_out& out, const uint32_t* arena);

// The seed loop and result assembly of the reference's seeded AlignOneWay (GraphAligner.h:408-491) given the
// per-seed direction traces.
AlignmentResult AssembleRead(const AlignmentGraph& graph, const ReadInput& read, const BatchPlan& plan, uint32_t readIndex,
	const std::vector<ga_stream_out>& outs, const std::vector<uint32_t>& arena);

// Implemented by the CUDA translation unit: runs all streams on the device behind ctx.
void ExecuteStreams(DeviceCtx* ctx, const std::vector<ga_stream_in>& streams, const std::vector<uint8_t>& parts, int initialBandwidth, int rampBandwidth,
	std::vector<ga_stream_out>& outs, std::vector<uint32_t>& arena, BatchStats* stats);

std::vector<AlignmentResult> AlignBatch(DeviceCtx* ctx, const AlignmentGraph& graph, const std::vector<ReadInput>& reads, int initialBandwidth, int rampBandwidth, BatchStats* stats);

}

#endif
