// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
//
// Driver around the UNMODIFIED reference hot path.  The reference's own translation units
// (GraphAlignerWrapper.cpp, AlignmentGraph.cpp, BigraphToDigraph.cpp, CommonUtils.cpp, ...)
// are compiled where they lie under /root/reference by oracle/Makefile (outputs only into
// oracle/_ref/).  This file only (a) parses a ".gacase" test case, (b) builds the
// AlignmentGraph through the reference's own DirectedGraph::Convert* + AddNode/AddEdgeNodeId/
// Finalize sequence (BigraphToDigraph.cpp:106-135 / :137-189), (c) calls the reference's
// seeded AlignOneWay (GraphAlignerWrapper.h:54) with the reference's thread model
// (Aligner.cpp:102-117: T std::threads popping a mutex-guarded LIFO), and (d) prints results.
//
// .gacase format (text, one record per line):
//   G vg | G gfa <overlap>
//   N <id> <sequence>
//   E <from> <from_start 0/1> <to> <to_end 0/1>
//   P <initialBandwidth> <rampBandwidth>
//   R <name> <sequence> <nseeds>      followed by <nseeds> lines:
//   S <bigraphNodeId> <readPos> <reverse 0/1>
//
// Output (stdout), per read in input order:
//   READ <name> failed=<0|1> score=<s> start=<a> end=<b> qpos=<q> nmap=<n> ntrace=<t> th=<fnv1a64 of trace>
//   M <digraph node_id> <is_reverse> <offset> <from_length> <to_length>         (one per mapping)
//   T <nodeID> <offset> <reverse> <readpos> <type>                              (only with --full)
// then one line: TIME threads=<T> reads=<n> aligned_bp=<bp> wall_ms=<ms>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <vector>
#include "AlignmentGraph.h"
#include "BigraphToDigraph.h"
#include "GraphAlignerWrapper.h"
#include "ThreadReadAssertion.h"

struct ReadCase
{
	std::string name;
	std::string sequence;
	std::vector<std::tuple<int, size_t, bool>> seeds;
};

struct ReadOut
{
	bool done = false;
	bool asserted = false;
	AlignmentResult result;
};

static uint64_t fnv(uint64_t h, uint64_t v)
{
	for (int i = 0; i < 8; i++)
	{
		h ^= (v >> (8 * i)) & 0xff;
		h *= 1099511628211ull;
	}
	return h;
}

int main(int argc, char** argv)
{
	if (argc < 2)
	{
		std::cerr << "usage: ref_align <case.gacase> [--threads T] [--full] [--quiet] [--limit N] [--summary]" << std::endl;
		return 2;
	}
	int threads = 1;
	bool full = false;
	bool quiet = false;
	bool summary = false;
	size_t limit = (size_t)-1;
	for (int i = 2; i < argc; i++)
	{
		if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--full")) full = true;
		else if (!strcmp(argv[i], "--quiet")) quiet = true;
		else if (!strcmp(argv[i], "--summary")) summary = true;
		else if (!strcmp(argv[i], "--limit") && i + 1 < argc) limit = strtoull(argv[++i], nullptr, 10);
	}
	std::ifstream in(argv[1]);
	if (!in.good())
	{
		std::cerr << "cannot open " << argv[1] << std::endl;
		return 2;
	}
	bool gfa = false;
	int overlap = 0;
	std::vector<std::pair<long, std::string>> nodes;
	std::vector<std::tuple<long, bool, long, bool>> edges;
	int initialBandwidth = 10, rampBandwidth = 0;
	std::vector<ReadCase> reads;
	std::string line;
	while (std::getline(in, line))
	{
		if (line.empty()) continue;
		std::stringstream ss(line);
		std::string tag;
		ss >> tag;
		if (tag == "G")
		{
			std::string kind;
			ss >> kind;
			gfa = (kind == "gfa");
			if (gfa) ss >> overlap;
		}
		else if (tag == "N")
		{
			long id;
			std::string seq;
			ss >> id >> seq;
			nodes.emplace_back(id, seq);
		}
		else if (tag == "E")
		{
			long from, to;
			int fs, te;
			ss >> from >> fs >> to >> te;
			edges.emplace_back(from, fs != 0, to, te != 0);
		}
		else if (tag == "P")
		{
			ss >> initialBandwidth >> rampBandwidth;
		}
		else if (tag == "R")
		{
			ReadCase r;
			size_t nseeds;
			ss >> r.name >> r.sequence >> nseeds;
			for (size_t i = 0; i < nseeds; i++)
			{
				std::getline(in, line);
				std::stringstream s2(line);
				std::string t2;
				int node;
				size_t pos;
				int rev;
				s2 >> t2 >> node >> pos >> rev;
				r.seeds.emplace_back(node, pos, rev != 0);
			}
			if (reads.size() < limit) reads.push_back(r);
		}
	}

	// silence the reference's stderr chatter (graph stats, "seed i/n ...") unless asked
	std::streambuf* oldcerr = nullptr;
	std::ofstream devnull("/dev/null");
	if (quiet) oldcerr = std::cerr.rdbuf(devnull.rdbuf());

	AlignmentGraph graph;
	graph.DBGOverlap = 0;
	if (gfa)
	{
		// BigraphToDigraph.cpp:137-189
		graph.DBGOverlap = overlap;
		for (auto& n : nodes)
		{
			std::string l = "S " + std::to_string(n.first) + " " + n.second;
			auto pair = DirectedGraph::ConvertGFANodeToNodes(l, graph.DBGOverlap);
			graph.AddNode(pair.first.nodeId, pair.first.sequence, !pair.first.rightEnd);
			graph.AddNode(pair.second.nodeId, pair.second.sequence, !pair.second.rightEnd);
		}
		for (auto& e : edges)
		{
			std::string l = "L " + std::to_string(std::get<0>(e)) + " " + (std::get<1>(e) ? "-" : "+") + " " + std::to_string(std::get<2>(e)) + " " + (std::get<3>(e) ? "-" : "+") + " " + std::to_string(overlap) + "M";
			auto pair = DirectedGraph::ConvertGFAEdgeToEdges(l);
			graph.AddEdgeNodeId(pair.first.fromId, pair.first.toId);
			graph.AddEdgeNodeId(pair.second.fromId, pair.second.toId);
		}
	}
	else
	{
		// BigraphToDigraph.cpp:106-135
		for (auto& n : nodes)
		{
			vg::Node vn;
			vn.set_id(n.first);
			vn.set_sequence(n.second);
			auto pair = DirectedGraph::ConvertVGNodeToNodes(vn);
			graph.AddNode(pair.first.nodeId, pair.first.sequence, !pair.first.rightEnd);
			graph.AddNode(pair.second.nodeId, pair.second.sequence, !pair.second.rightEnd);
		}
		for (auto& e : edges)
		{
			vg::Edge ve;
			ve.set_from(std::get<0>(e));
			ve.set_from_start(std::get<1>(e));
			ve.set_to(std::get<2>(e));
			ve.set_to_end(std::get<3>(e));
			auto pair = DirectedGraph::ConvertVGEdgeToEdges(ve);
			graph.AddEdgeNodeId(pair.first.fromId, pair.first.toId);
			graph.AddEdgeNodeId(pair.second.fromId, pair.second.toId);
		}
	}
	graph.Finalize(64);

	std::vector<ReadOut> outs(reads.size());
	std::vector<size_t> stack;
	for (size_t i = 0; i < reads.size(); i++) stack.push_back(i);
	std::mutex mutex;
	auto worker = [&]() {
		while (true)
		{
			size_t idx;
			{
				std::lock_guard<std::mutex> guard(mutex);
				if (stack.empty()) return;
				idx = stack.back();
				stack.pop_back();
			}
			const ReadCase& r = reads[idx];
			ThreadReadAssertion::setRead(r.name);
			try
			{
				if (r.seeds.empty())
				{
					outs[idx].result.alignmentFailed = true; // Aligner.cpp:131-138 "has no seed hits"
				}
				else
				{
					outs[idx].result = AlignOneWay(graph, r.name, r.sequence, initialBandwidth, rampBandwidth, 64, r.seeds);
				}
			}
			catch (const ThreadReadAssertion::AssertionFailure&)
			{
				outs[idx].asserted = true; // Aligner.cpp:143-148
				outs[idx].result.alignmentFailed = true;
			}
			outs[idx].done = true;
		}
	};
	auto t0 = std::chrono::steady_clock::now();
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; t++) pool.emplace_back(worker);
	for (auto& t : pool) t.join();
	auto t1 = std::chrono::steady_clock::now();
	if (quiet) std::cerr.rdbuf(oldcerr);

	size_t alignedBp = 0;
	for (size_t i = 0; i < reads.size(); i++)
	{
		const AlignmentResult& r = outs[i].result;
		bool failed = r.alignmentFailed;
		uint64_t h = 14695981039346656037ull;
		for (auto& t : r.trace)
		{
			h = fnv(h, (uint64_t)(int64_t)t.nodeID);
			h = fnv(h, t.offset);
			h = fnv(h, t.reverse ? 1 : 0);
			h = fnv(h, t.readpos);
			h = fnv(h, (uint64_t)t.type);
		}
		if (!failed) alignedBp += reads[i].sequence.size();
		int nmap = failed ? 0 : r.alignment.path().mapping_size();
		if (summary)
		{
			// --summary: the mappings as one order-sensitive checksum instead of a line each (full-size read sets):
			// mh = sum over mappings m of (m + 1) * mix(m) mod 2^64, see tools/gacase.py mapping_checksums
			uint64_t mh = 0;
			for (int m = 0; m < nmap; m++)
			{
				const auto& mp = r.alignment.path().mapping(m);
				const uint64_t mix = (uint64_t)(int64_t)mp.position().node_id() * 0x9E3779B97F4A7C15ull + (uint64_t)(mp.position().is_reverse() ? 1 : 0) * 0xC2B2AE3D27D4EB4Full
					+ (uint64_t)(int64_t)mp.position().offset() * 0x165667B19E3779F9ull + (uint64_t)(int64_t)(mp.edit_size() > 0 ? mp.edit(0).from_length() : -1) * 0x27D4EB2F165667C5ull
					+ (uint64_t)(int64_t)(mp.edit_size() > 0 ? mp.edit(0).to_length() : -1) * 0x85EBCA77C2B2AE63ull;
				mh += (uint64_t)(m + 1) * mix;
			}
			printf("READ %s failed=%d asserted=%d score=%d start=%zu end=%zu qpos=%d nmap=%d ntrace=%zu th=%016llx mh=%016llx\n", reads[i].name.c_str(), failed ? 1 : 0, outs[i].asserted ? 1 : 0, failed ? 0 : r.alignment.score(), failed ? (size_t)0 : r.alignmentStart, failed ? (size_t)0 : r.alignmentEnd, failed ? 0 : r.alignment.query_position(), nmap, failed ? (size_t)0 : r.trace.size(), (unsigned long long)(failed ? 0 : h), (unsigned long long)mh);
			continue;
		}
		printf("READ %s failed=%d asserted=%d score=%d start=%zu end=%zu qpos=%d nmap=%d ntrace=%zu th=%016llx\n", reads[i].name.c_str(), failed ? 1 : 0, outs[i].asserted ? 1 : 0, failed ? 0 : r.alignment.score(), failed ? (size_t)0 : r.alignmentStart, failed ? (size_t)0 : r.alignmentEnd, failed ? 0 : r.alignment.query_position(), nmap, failed ? (size_t)0 : r.trace.size(), (unsigned long long)(failed ? 0 : h));
		for (int m = 0; m < nmap; m++)
		{
			const auto& mp = r.alignment.path().mapping(m);
			printf("M %lld %d %lld %d %d\n", (long long)mp.position().node_id(), mp.position().is_reverse() ? 1 : 0, (long long)mp.position().offset(), mp.edit_size() > 0 ? mp.edit(0).from_length() : -1, mp.edit_size() > 0 ? mp.edit(0).to_length() : -1);
		}
		if (full && !failed)
		{
			for (auto& t : r.trace)
			{
				printf("T %d %zu %d %zu %d\n", t.nodeID, t.offset, t.reverse ? 1 : 0, t.readpos, (int)t.type);
			}
		}
	}
	double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
	printf("TIME threads=%d reads=%zu aligned_bp=%zu wall_ms=%.3f\n", threads, reads.size(), alignedBp, ms);
	return 0;
}
