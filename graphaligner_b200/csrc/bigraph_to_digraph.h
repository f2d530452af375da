// Bidirected (vg / GFA) graph -> directed AlignmentGraph: every node id becomes two digraph nodes, 2*id
// (forward) and 2*id+1 (reverse complement), every edge two directed edges (reference
// BigraphToDigraph.cpp:27-104).  Same call order as the reference loaders (all nodes, then all edges) so that
// node indices and neighbour order - and with them the traceback tie-breaks - are identical.
#ifndef GA_BIGRAPH_TO_DIGRAPH_H
#define GA_BIGRAPH_TO_DIGRAPH_H
#include <cstdint>
#include <string>
#include <vector>
#include "alignment_graph.h"

namespace DirectedGraph
{
struct BiNode
{
	int64_t id;
	std::string sequence;
	std::string name;   // vg Node.name: carried through to the -A augmented graph only
};
struct BiEdge
{
	int64_t from;
	int64_t to;
	bool from_start;
	bool to_end;
};
// vg semantics (StreamVGGraphFromFile, BigraphToDigraph.cpp:106-135)
AlignmentGraph BuildFromVG(const std::vector<BiNode>& nodes, const std::vector<BiEdge>& edges);
// GFA semantics: node sequences lose their last `overlap` bases on both strands, DBGOverlap = overlap
// (StreamGFAGraphFromFile, BigraphToDigraph.cpp:137-189; '-' orientation == from_start / to_end)
AlignmentGraph BuildFromGFA(const std::vector<BiNode>& nodes, const std::vector<BiEdge>& edges, int overlap);
AlignmentGraph StreamVGGraphFromFile(const std::string& filename);
AlignmentGraph StreamGFAGraphFromFile(const std::string& filename);
}

#endif
