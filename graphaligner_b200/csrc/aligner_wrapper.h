// Drop-in for the reference's GraphAlignerWrapper.h: the same two free functions with the same signatures
// (reference GraphAlignerWrapper.h:53-54), returning the same AlignmentResult, computed on the GPU.
// A process-wide engine (one device context + graph replica per AlignmentGraph object) is created on first use.
#ifndef GA_ALIGNER_WRAPPER_H
#define GA_ALIGNER_WRAPPER_H
#include <string>
#include <tuple>
#include <vector>
#include "alignment_graph.h"
#include "ga_host.h"

// seeded alignment - the working entry point of the reference (GraphAligner.h:408-491)
AlignmentResult AlignOneWay(const AlignmentGraph& graph, const std::string& seq_id, const std::string& sequence, int initialBandwidth, int rampBandwidth, size_t dynamicRowStart,
	const std::vector<std::tuple<int, size_t, bool>>& seedHits);
// full-band start without seeds: asserts immediately in the reference at this commit (GraphAligner.h:1138, SURVEY.md
// quirk 1); here it reports a failed alignment instead of computing something the reference never could
AlignmentResult AlignOneWay(const AlignmentGraph& graph, const std::string& seq_id, const std::string& sequence, int initialBandwidth, int rampBandwidth, size_t dynamicRowStart);

struct AlignerRead
{
	std::string name;
	std::string sequence;
	std::vector<std::tuple<int, size_t, bool>> seedHits;
};
// the batched form the driver should prefer: one GPU launch for all reads
std::vector<AlignmentResult> AlignReads(const AlignmentGraph& graph, const std::vector<AlignerRead>& reads, int initialBandwidth, int rampBandwidth, int device = 0);
// the same over several GPUs of the node (the counterpart of the reference's worker threads, Aligner.cpp:285-306): the read
// set is cut into batches that the devices pull from one queue; results come back in input order
std::vector<AlignmentResult> AlignReads(const AlignmentGraph& graph, const std::vector<AlignerRead>& reads, int initialBandwidth, int rampBandwidth, const std::vector<int>& devices);
// drops the cached engine of a graph (call before destroying the graph)
void ReleaseAlignerEngine(const AlignmentGraph& graph);

#endif
