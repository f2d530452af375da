// Host-side sequence graph with the reference's AlignmentGraph surface (reference AlignmentGraph.h:25-43),
// stored the way the GPU wants it: CSR adjacency (insertion order kept, duplicates dropped), 64-bit node
// offsets and 2-bit packed bases.  Finalize() produces exactly the arrays ga_graph_view points at.
#ifndef GA_ALIGNMENT_GRAPH_H
#define GA_ALIGNMENT_GRAPH_H
#include <cstddef>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include "ga_types.h"

class AlignmentGraph
{
public:
	AlignmentGraph();
	void ReserveNodes(size_t numNodes, size_t totalSequenceLength);
	// nodeId is the digraph id (2*id forward, 2*id+1 reverse complement, BigraphToDigraph.cpp:27-30)
	void AddNode(int nodeId, const std::string& sequence, bool reverseNode);
	void AddEdgeNodeId(int node_id_from, int node_id_to);
	void Finalize(int wordSize);
	bool Finalized() const { return finalized; }

	size_t GetReverseNode(size_t nodeIndex) const;
	size_t GetReversePosition(size_t position) const;
	size_t SizeInBp() const { return totalBp; }
	size_t IndexToNode(size_t index) const;
	size_t NodeSize() const { return nodeIDs.size(); }
	size_t NodeStart(size_t nodeIndex) const { return nodeStart[nodeIndex]; }
	size_t NodeEnd(size_t nodeIndex) const { return nodeStart[nodeIndex + 1]; }
	size_t NodeLength(size_t nodeIndex) const { return nodeStart[nodeIndex + 1] - nodeStart[nodeIndex]; }
	char NodeSequences(size_t index) const;
	size_t NodeSequencesSize() const { return totalBp; }
	int NodeID(size_t nodeIndex) const { return nodeIDs[nodeIndex]; }
	bool Reverse(size_t nodeIndex) const { return reverse[nodeIndex] != 0; }
	// throws std::out_of_range for an unknown id, like nodeLookup.at (GraphAligner.h:423)
	size_t Lookup(int digraphNodeId) const;
	bool HasNode(int digraphNodeId) const;
	size_t DummyNodeStart() const { return 0; }
	size_t DummyNodeEnd() const { return nodeIDs.size() - 1; }
	size_t NumEdges() const { return inAdj.size(); }
	bool HasOutNeighbor(size_t from, size_t to) const;

	// flat arrays for the device (valid after Finalize)
	ga_graph_view View() const;
	const std::vector<uint64_t>& NodeStarts() const { return nodeStart; }
	const std::vector<uint32_t>& Seq2() const { return seq2; }
	const std::vector<uint32_t>& InOff() const { return inOff; }
	const std::vector<uint32_t>& InAdj() const { return inAdj; }
	const std::vector<uint32_t>& OutOff() const { return outOff; }
	const std::vector<uint32_t>& OutAdj() const { return outAdj; }

	int DBGOverlap;
	// distinguishes graph objects beyond their address (a cached device replica must not outlive its graph's identity)
	uint64_t Uid() const { return uid; }

private:
	void pushBase(unsigned code);
	std::vector<uint64_t> nodeStart;   // while building: one entry per node; Finalize appends the total
	std::vector<int> nodeIDs;
	std::vector<uint8_t> reverse;
	std::vector<uint32_t> reverseNode;   // index of the node on the other strand (0xffffffff: none), built by Finalize
	std::vector<uint32_t> seq2;
	uint64_t totalBp;
	std::unordered_map<int, uint32_t> nodeLookup;
	std::vector<uint32_t> denseLookup;   // digraph id -> node index (0xffffffff: none) when the ids are dense, built by Finalize; else empty
	std::vector<std::pair<uint32_t, uint32_t>> pendingEdges;   // (from, to) in AddEdgeNodeId order
	std::vector<uint32_t> inOff, inAdj, outOff, outAdj;
	bool finalized;
	uint64_t uid = 0;    // assigned by Finalize
};

#endif
