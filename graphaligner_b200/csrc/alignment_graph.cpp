#include "alignment_graph.h"
#include <atomic>
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

// Base codes used on the device: A=0, C=1, G=2, T=3.  (The reference keeps two vector<bool> planes,
// AlignmentGraph.cpp:61-85; only A/C/G/T are legal node characters there too.)
static int baseCode(char c)
{
	switch (c)
	{
		case 'A': return 0;
		case 'C': return 1;
		case 'G': return 2;
		case 'T': return 3;
		default: return -1;
	}
}

AlignmentGraph::AlignmentGraph() : DBGOverlap(0), totalBp(0), finalized(false)
{
	// dummy start node: index 0, 1 bp, no edges (AlignmentGraph.cpp:22-30)
	nodeIDs.push_back(0);
	nodeStart.push_back(0);
	reverse.push_back(0);
	pushBase(0);
}

void AlignmentGraph::pushBase(unsigned code)
{
	if ((totalBp & 15) == 0) seq2.push_back(0);
	seq2.back() |= code << ((totalBp & 15) * 2);
	totalBp++;
}

void AlignmentGraph::ReserveNodes(size_t numNodes, size_t totalSequenceLength)
{
	nodeIDs.reserve(numNodes + 2);
	nodeStart.reserve(numNodes + 3);
	reverse.reserve(numNodes + 2);
	seq2.reserve((totalSequenceLength + 2) / 16 + 1);
	nodeLookup.reserve(numNodes + 2);
}

void AlignmentGraph::AddNode(int nodeId, const std::string& sequence, bool reverseNode)
{
	if (finalized) throw std::logic_error("AlignmentGraph::AddNode after Finalize");
	// duplicate ids are ignored (AlignmentGraph.cpp:49-51)
	if (nodeLookup.count(nodeId) != 0) return;
	for (char c : sequence)
	{
		if (baseCode(c) < 0)
		{
			// the reference aborts on anything but ACGT (AlignmentGraph.cpp:80-83)
			std::cerr << "AlignmentGraph: illegal character '" << c << "' in node " << nodeId << std::endl;
			std::abort();
		}
	}
	nodeLookup[nodeId] = (uint32_t)nodeIDs.size();
	nodeIDs.push_back(nodeId);
	nodeStart.push_back(totalBp);
	reverse.push_back(reverseNode ? 1 : 0);
	for (char c : sequence) pushBase((unsigned)baseCode(c));
}

void AlignmentGraph::AddEdgeNodeId(int node_id_from, int node_id_to)
{
	if (finalized) throw std::logic_error("AlignmentGraph::AddEdgeNodeId after Finalize");
	auto f = nodeLookup.find(node_id_from);
	auto t = nodeLookup.find(node_id_to);
	if (f == nodeLookup.end() || t == nodeLookup.end()) throw std::out_of_range("AlignmentGraph::AddEdgeNodeId: unknown node id");
	pendingEdges.emplace_back(f->second, t->second);
}

// Stable counting sort of the edge list into CSR, then per-node removal of duplicates keeping the first
// occurrence: the same lists, in the same order, as the reference's push_back-unless-found
// (AlignmentGraph.cpp:104-105).
static void buildCsr(size_t n, const std::vector<std::pair<uint32_t, uint32_t>>& edges, bool byTarget, std::vector<uint32_t>& off, std::vector<uint32_t>& adj)
{
	std::vector<uint32_t> count(n + 1, 0);
	for (auto& e : edges) count[(byTarget ? e.second : e.first) + 1]++;
	for (size_t i = 0; i < n; i++) count[i + 1] += count[i];
	std::vector<uint32_t> raw(edges.size());
	std::vector<uint32_t> fill(count.begin(), count.end() - 1);
	for (auto& e : edges)
	{
		uint32_t key = byTarget ? e.second : e.first;
		raw[fill[key]++] = byTarget ? e.first : e.second;
	}
	off.assign(n + 1, 0);
	adj.clear();
	adj.reserve(edges.size());
	for (size_t i = 0; i < n; i++)
	{
		off[i] = (uint32_t)adj.size();
		size_t begin = adj.size();
		for (uint32_t k = count[i]; k < count[i + 1]; k++)
		{
			bool dup = false;
			for (size_t q = begin; q < adj.size(); q++)
			{
				if (adj[q] == raw[k]) { dup = true; break; }
			}
			if (!dup) adj.push_back(raw[k]);
		}
	}
	off[n] = (uint32_t)adj.size();
	adj.shrink_to_fit();
}

void AlignmentGraph::Finalize(int wordSize)
{
	(void)wordSize;
	if (finalized) return;
	// dummy end node (AlignmentGraph.cpp:108-118)
	nodeIDs.push_back(0);
	nodeStart.push_back(totalBp);
	reverse.push_back(0);
	pushBase(0);
	nodeStart.push_back(totalBp);
	size_t n = nodeIDs.size();
	buildCsr(n, pendingEdges, true, inOff, inAdj);
	buildCsr(n, pendingEdges, false, outOff, outAdj);
	pendingEdges.clear();
	pendingEdges.shrink_to_fit();
	// keep one spare word so a 16-base window read at the very end stays in bounds
	seq2.push_back(0);   // two words of padding: the device reads 32 bases (three words) at a time
	seq2.push_back(0);
	size_t special = 0;
	for (size_t i = 0; i < n; i++)
	{
		if (inOff[i + 1] - inOff[i] >= 2) special++;
	}
	static std::atomic<uint64_t> nextUid(1);
	uid = nextUid.fetch_add(1);
	reverseNode.assign(n, 0xffffffffu);
	for (size_t i = 0; i < n; i++)
	{
		const int id = nodeIDs[i];
		auto found = nodeLookup.find(id % 2 == 1 ? (id / 2) * 2 : (id / 2) * 2 + 1);
		if (found != nodeLookup.end()) reverseNode[i] = found->second;
	}
	// digraph id -> node index as a flat table when the ids are dense (every seed costs several lookups: a hash probe is two
	// or three cache misses, the table one)
	{
		// from the map itself: the two dummy nodes carry id 0 in nodeIDs without being looked up by it
		int maxId = -1, minId = 0;
		for (const auto& kv : nodeLookup) { maxId = std::max(maxId, kv.first); minId = std::min(minId, kv.first); }
		denseLookup.clear();
		if (minId >= 0 && maxId >= 0 && (size_t)maxId < 4 * n + 1024)
		{
			denseLookup.assign((size_t)maxId + 1, 0xffffffffu);
			for (const auto& kv : nodeLookup) denseLookup[(size_t)kv.first] = kv.second;
		}
	}
	// same graph statistics on stderr as the reference (AlignmentGraph.cpp:125-138)
	std::cerr << n << " nodes" << std::endl;
	std::cerr << totalBp << "bp" << std::endl;
	std::cerr << inAdj.size() << " edges" << std::endl;
	std::cerr << special << " nodes with in-degree >= 2" << std::endl;
	finalized = true;
}

size_t AlignmentGraph::Lookup(int digraphNodeId) const
{
	if (!denseLookup.empty())
	{
		if (digraphNodeId < 0 || (size_t)digraphNodeId >= denseLookup.size() || denseLookup[(size_t)digraphNodeId] == 0xffffffffu) throw std::out_of_range("AlignmentGraph: node id not in graph");
		return denseLookup[(size_t)digraphNodeId];
	}
	auto found = nodeLookup.find(digraphNodeId);
	if (found == nodeLookup.end()) throw std::out_of_range("AlignmentGraph: node id not in graph");
	return found->second;
}

bool AlignmentGraph::HasNode(int digraphNodeId) const
{
	if (!denseLookup.empty()) return digraphNodeId >= 0 && (size_t)digraphNodeId < denseLookup.size() && denseLookup[(size_t)digraphNodeId] != 0xffffffffu;
	return nodeLookup.count(digraphNodeId) != 0;
}

size_t AlignmentGraph::GetReverseNode(size_t nodeIndex) const
{
	// AlignmentGraph.cpp:199-214; the finalized graph answers from a table (one lookup per run of every trace)
	if (finalized && reverseNode[nodeIndex] != 0xffffffffu) return reverseNode[nodeIndex];
	int id = nodeIDs[nodeIndex];
	int bigraphNodeId = id / 2;
	return Lookup(id % 2 == 1 ? bigraphNodeId * 2 : bigraphNodeId * 2 + 1);
}

size_t AlignmentGraph::GetReversePosition(size_t pos) const
{
	// AlignmentGraph.cpp:216-224
	size_t originalNode = IndexToNode(pos);
	size_t otherNode = GetReverseNode(originalNode);
	return (NodeEnd(otherNode) - 1) - (pos - nodeStart[originalNode]);
}

size_t AlignmentGraph::IndexToNode(size_t index) const
{
	// AlignmentGraph.cpp:226-234; nodeStart carries one extra entry (the total) after Finalize
	auto next = std::upper_bound(nodeStart.begin(), nodeStart.begin() + nodeIDs.size(), (uint64_t)index);
	return (size_t)(next - nodeStart.begin()) - 1;
}

char AlignmentGraph::NodeSequences(size_t index) const
{
	if (index == 0 || index == totalBp - 1) return '-';
	return "ACGT"[(seq2[index >> 4] >> ((index & 15) * 2)) & 3];
}

bool AlignmentGraph::HasOutNeighbor(size_t from, size_t to) const
{
	for (uint32_t e = outOff[from]; e < outOff[from + 1]; e++)
	{
		if (outAdj[e] == to) return true;
	}
	return false;
}

ga_graph_view AlignmentGraph::View() const
{
	ga_graph_view v;
	v.nNodes = (uint32_t)nodeIDs.size();
	v.nodeStart = nodeStart.data();
	v.seq2 = seq2.data();
	v.inOff = inOff.data();
	v.inAdj = inAdj.data();
	v.outOff = outOff.data();
	v.outAdj = outAdj.data();
	v.nodeRec = nullptr;
	v.seqChunks = nullptr;
	v.nodeIdRev = nullptr;
	return v;
}
