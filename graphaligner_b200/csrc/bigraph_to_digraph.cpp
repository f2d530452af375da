#include "bigraph_to_digraph.h"
#include <fstream>
#include <sstream>
#include <stdexcept>
#include "ga_host.h"
#include "vg_codec.h"

namespace DirectedGraph
{

static void addEdgePair(AlignmentGraph& g, const BiEdge& e)
{
	// "right" = the side a walk leaves a node from, "left" = the side it enters; an edge a->b also allows the
	// reverse-complement walk b'->a' (BigraphToDigraph.cpp:32-56)
	int fromRight = (int)(e.from * 2 + (e.from_start ? 1 : 0));
	int fromLeft = (int)(e.from * 2 + (e.from_start ? 0 : 1));
	int toRight = (int)(e.to * 2 + (e.to_end ? 1 : 0));
	int toLeft = (int)(e.to * 2 + (e.to_end ? 0 : 1));
	g.AddEdgeNodeId(fromRight, toRight);
	g.AddEdgeNodeId(toLeft, fromLeft);
}

AlignmentGraph BuildFromVG(const std::vector<BiNode>& nodes, const std::vector<BiEdge>& edges)
{
	AlignmentGraph g;
	size_t bp = 0;
	for (auto& n : nodes) bp += n.sequence.size();
	g.ReserveNodes(nodes.size() * 2, bp * 2);
	for (auto& n : nodes)
	{
		g.AddNode((int)(n.id * 2), n.sequence, false);
		g.AddNode((int)(n.id * 2 + 1), ga::ReverseComplement(n.sequence), true);
	}
	for (auto& e : edges) addEdgePair(g, e);
	g.Finalize(64);
	return g;
}

AlignmentGraph BuildFromGFA(const std::vector<BiNode>& nodes, const std::vector<BiEdge>& edges, int overlap)
{
	AlignmentGraph g;
	g.DBGOverlap = overlap;
	size_t bp = 0;
	for (auto& n : nodes) bp += n.sequence.size();
	g.ReserveNodes(nodes.size() * 2, bp * 2);
	for (auto& n : nodes)
	{
		if ((int)n.sequence.size() <= overlap) throw std::runtime_error("GFA node shorter than the edge overlap");
		size_t keep = n.sequence.size() - overlap;
		g.AddNode((int)(n.id * 2), n.sequence.substr(0, keep), false);
		g.AddNode((int)(n.id * 2 + 1), ga::ReverseComplement(n.sequence).substr(0, keep), true);
	}
	for (auto& e : edges) addEdgePair(g, e);
	g.Finalize(64);
	return g;
}

AlignmentGraph StreamVGGraphFromFile(const std::string& filename)
{
	std::vector<BiNode> nodes;
	std::vector<BiEdge> edges;
	vgcodec::ReadGraphFile(filename, nodes, edges);
	return BuildFromVG(nodes, edges);
}

AlignmentGraph StreamGFAGraphFromFile(const std::string& filename)
{
	std::vector<BiNode> nodes;
	std::vector<BiEdge> edges;
	int overlap = 0;
	std::ifstream file(filename);
	std::string line;
	while (std::getline(file, line))
	{
		if (line.empty()) continue;
		std::stringstream str(line);
		std::string tag;
		str >> tag;
		if (tag == "S")
		{
			BiNode n;
			str >> n.id >> n.sequence;
			nodes.push_back(n);
		}
		else if (tag == "L")
		{
			BiEdge e;
			std::string fromOrient, toOrient, ov;
			str >> e.from >> fromOrient >> e.to >> toOrient >> ov;
			e.from_start = fromOrient == "-";
			e.to_end = toOrient == "-";
			edges.push_back(e);
			if (!ov.empty()) overlap = std::stoi(ov.substr(0, ov.size() - 1));   // "<n>M", last L line wins like the reference
		}
	}
	return BuildFromGFA(nodes, edges, overlap);
}

}
