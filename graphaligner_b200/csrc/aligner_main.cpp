// bin/Aligner: the reference's command line (AlignerMain.cpp:8-101) and driver (Aligner.cpp:231-322) over the GPU hot path.
//   -g graph (.vg / .gfa)   -f reads (.fastq/.fq/.fasta/.fa)   -s seeds (GAM)   -a output GAM   -t threads
//   -b initial bandwidth    -B ramp bandwidth   -d dynamic row start (multiple of 64, unused like upstream)   -i (dead upstream)
//   -A augmented graph (.vg input only, as upstream)   -G cuda devices (extra: "0", "0-7", "0,2")
// Same validation messages and exit(0) paths; same per-read log lines, alignment_<t>_<read>.gam and trace_<t>_<read>.trace files.
// Differences: all reads are aligned in GPU batches, reported as "thread 0"; -t only sizes the host worker pool.
#include <unistd.h>
#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <limits>
#include <atomic>
#include <thread>
#include <map>
#include <string>
#include <vector>
#include "aligner_wrapper.h"
#include "bigraph_to_digraph.h"
#include "vg_codec.h"

struct FastQ
{
	std::string seq_id, sequence;
};

// fastqloader.cpp:6-70: format by extension; 4-line FASTQ records / multi-line FASTA
static std::vector<FastQ> loadFastqFromFile(const std::string& filename)
{
	auto ends = [&](const char* suffix) { std::string s(suffix); return filename.size() >= s.size() && filename.substr(filename.size() - s.size()) == s; };
	std::vector<FastQ> result;
	std::ifstream file(filename);
	std::string line;
	auto chomp = [](std::string& l) { if (!l.empty() && l.back() == '\r') l.pop_back(); };
	if (ends(".fastq") || ends(".fq"))
	{
		while (std::getline(file, line))
		{
			if (line.empty() || line[0] != '@') continue;
			chomp(line);
			FastQ r;
			r.seq_id = line.substr(1);
			std::getline(file, line);
			chomp(line);
			r.sequence = line;
			std::getline(file, line);
			std::getline(file, line);
			result.push_back(r);
		}
	}
	else if (ends(".fasta") || ends(".fa"))
	{
		FastQ cur;
		bool have = false;
		while (std::getline(file, line))
		{
			chomp(line);
			if (!line.empty() && line[0] == '>')
			{
				if (have) result.push_back(cur);
				cur = FastQ();
				cur.seq_id = line.substr(1);
				have = true;
			}
			else if (have) cur.sequence += line;
		}
		if (have) result.push_back(cur);
	}
	return result;
}

static bool fileExists(const std::string& name)
{
	std::ifstream f(name);
	return f.good();
}

static std::string sanitize(std::string name)
{
	std::replace(name.begin(), name.end(), '/', '_');
	std::replace(name.begin(), name.end(), ':', '_');
	return name;
}

// -G: the GPUs to align on: "3", "0,2,5" or "0-7" (not in the reference: its parallelism is -t host threads)
static std::vector<int> parseDevices(const std::string& arg)
{
	std::vector<int> devices;
	size_t pos = 0;
	while (pos <= arg.size())
	{
		size_t comma = arg.find(',', pos);
		if (comma == std::string::npos) comma = arg.size();
		const std::string item = arg.substr(pos, comma - pos);
		const size_t dash = item.find('-');
		if (dash != std::string::npos && dash > 0)
		{
			const int lo = std::stoi(item.substr(0, dash)), hi = std::stoi(item.substr(dash + 1));
			for (int d = lo; d <= hi; d++) devices.push_back(d);
		}
		else if (!item.empty()) devices.push_back(std::stoi(item));
		pos = comma + 1;
	}
	return devices;
}

// augmentGraphwithAlignment, Aligner.cpp:24-74: the input graph's nodes, and one edge per pair of consecutive mappings of
// every alignment (node ids already back to the original ones)
static void writeAugmentedGraph(const std::string& graphFile, const std::string& outFile, const std::vector<vg::Alignment>& alignments)
{
	std::vector<DirectedGraph::BiNode> nodes;
	std::vector<DirectedGraph::BiEdge> inputEdges, edges;
	vgcodec::ReadGraphFile(graphFile, nodes, inputEdges);
	for (auto& a : alignments)
	{
		for (size_t i = 0; i + 1 < a.path.mapping.size(); i++)
		{
			const auto& from = a.path.mapping[i].position;
			const auto& to = a.path.mapping[i + 1].position;
			edges.push_back(DirectedGraph::BiEdge { (int64_t)from.node_id, (int64_t)to.node_id, from.is_reverse, to.is_reverse });
		}
	}
	vgcodec::WriteStreamFile(outFile, { vgcodec::EncodeGraph(nodes, edges) });
}

int main(int argc, char** argv)
{
	std::string graphFile, fastqFile, alignmentFile, auggraphFile, seedFile;
	int numThreads = 0, initialBandwidth = 0, rampBandwidth = 0, dynamicRowStart = 64;
	std::vector<int> devices;
	bool initialFullBand = false;
	int c;
	while ((c = getopt(argc, argv, "g:f:a:t:B:A:is:d:MSb:G:")) != -1)
	{
		switch (c)
		{
			case 'g': graphFile = optarg; break;
			case 'f': fastqFile = optarg; break;
			case 'a': alignmentFile = optarg; break;
			case 't': numThreads = std::stoi(optarg); break;
			case 'b': initialBandwidth = std::stoi(optarg); break;
			case 'B': rampBandwidth = std::stoi(optarg); break;
			case 'A': auggraphFile = optarg; break;
			case 'i': initialFullBand = true; break;
			case 's': seedFile = optarg; break;
			case 'd': dynamicRowStart = std::stoi(optarg); break;
			case 'G': devices = parseDevices(optarg); break;
		}
	}
	// AlignerMain.cpp:68-96
	if (dynamicRowStart % 64 != 0) { std::cerr << "dynamic row start has to be a multiple of 64" << std::endl; std::exit(0); }
	if (numThreads < 1) { std::cerr << "number of threads must be >= 1" << std::endl; std::exit(0); }
	if (initialBandwidth < 2) { std::cerr << "bandwidth must be >= 2" << std::endl; std::exit(0); }
	if (rampBandwidth != 0 && rampBandwidth <= initialBandwidth) { std::cerr << "backup bandwidth must be higher than initial bandwidth" << std::endl; std::exit(0); }
	if (!initialFullBand && seedFile == "") { std::cerr << "either initial full band or seed file must be set" << std::endl; std::exit(0); }
	setenv("GA_HOST_THREADS", std::to_string(numThreads).c_str(), 0);

	// alignReads, Aligner.cpp:231-322
	std::vector<FastQ> fastqs;
	if (fileExists(fastqFile))
	{
		fastqs = loadFastqFromFile(fastqFile);
		std::cout << fastqs.size() << " reads" << std::endl;
	}
	else { std::cerr << "No fastq file exists" << std::endl; std::exit(0); }

	std::map<std::string, std::vector<std::tuple<int, size_t, bool>>> seeds;
	bool haveSeeds = seedFile != "";
	if (haveSeeds)
	{
		if (!fileExists(seedFile)) { std::cerr << "No seeds file exists" << std::endl; std::exit(0); }
		for (auto& a : vgcodec::ReadAlignmentFile(seedFile))
		{
			if (a.path.mapping.empty()) continue;
			seeds[a.name].emplace_back((int)a.path.mapping[0].position.node_id, (size_t)a.query_position, a.path.mapping[0].position.is_reverse);
		}
	}
	if (fileExists(graphFile)) std::cout << "load graph from " << graphFile << std::endl;
	else { std::cerr << "No graph file exists" << std::endl; std::exit(0); }
	AlignmentGraph graph;
	if (graphFile.size() >= 3 && graphFile.substr(graphFile.size() - 3) == ".vg") graph = DirectedGraph::StreamVGGraphFromFile(graphFile);
	else if (graphFile.size() >= 4 && graphFile.substr(graphFile.size() - 4) == ".gfa") graph = DirectedGraph::StreamGFAGraphFromFile(graphFile);
	else { std::cerr << "Unknown graph type (" << graphFile << ")" << std::endl; std::exit(0); }

	// the reference pops reads from the back of the list (Aligner.cpp:111-117): last read first
	std::vector<AlignerRead> batch;
	std::vector<size_t> batchRead;
	for (size_t k = fastqs.size(); k-- > 0;)
	{
		AlignerRead r;
		r.name = fastqs[k].seq_id;
		r.sequence = fastqs[k].sequence;
		auto found = seeds.find(r.name);
		if (found != seeds.end()) r.seedHits = found->second;
		// a seed on a node that is not in the graph terminates the reference (std::out_of_range); here the read fails
		batch.push_back(std::move(r));
		batchRead.push_back(k);
	}
	std::vector<AlignmentResult> results;
	if (initialFullBand && !haveSeeds)
	{
		// -i without seeds asserts at once upstream (GraphAligner.h:1138): every read fails
		results.resize(batch.size());
		for (auto& r : results) r.alignment.score = std::numeric_limits<int32_t>::max();
	}
	else if (getenv("GA_PER_READ"))
	{
		// the reference's own structure (Aligner.cpp:107-140,285-298): -t worker threads, each calling the per-read
		// AlignOneWay; kept for checking the drop-in entry point, the batched call below is what the GPU wants
		results.resize(batch.size());
		std::atomic<size_t> next(0);
		std::vector<std::thread> workers;
		for (int t = 0; t < numThreads; t++)
		{
			workers.emplace_back([&]() {
				while (true)
				{
					const size_t i = next.fetch_add(1);
					if (i >= batch.size()) break;
					if (haveSeeds && batch[i].seedHits.empty()) { results[i].alignment.score = std::numeric_limits<int32_t>::max(); continue; }
					try
					{
						results[i] = AlignOneWay(graph, batch[i].name, batch[i].sequence, initialBandwidth, rampBandwidth, (size_t)dynamicRowStart, batch[i].seedHits);
					}
					catch (const std::out_of_range&)
					{
						results[i] = AlignmentResult();
						results[i].alignmentFailed = true;
						results[i].alignment.score = std::numeric_limits<int32_t>::max();
					}
				}
			});
		}
		for (auto& w : workers) w.join();
	}
	else
	{
		results = AlignReads(graph, batch, initialBandwidth, rampBandwidth, devices);
	}
	std::vector<vg::Alignment> alignments;
	const int threadnum = 0;
	for (size_t i = 0; i < batch.size(); i++)
	{
		const AlignerRead& read = batch[i];
		AlignmentResult& alignment = results[i];
		std::cout << "thread " << threadnum << " " << (batch.size() - 1 - i) << " left\n";
		std::cout << "read " << read.name << " size " << read.sequence.size() << "bp" << std::endl;
		if (haveSeeds && read.seedHits.empty())
		{
			std::cout << "read " << read.name << " has no seed hits" << std::endl;
			std::cerr << "read " << read.name << " has no seed hits" << std::endl;
			std::cout << "read " << read.name << " alignment failed" << std::endl;
			std::cerr << "read " << read.name << " alignment failed" << std::endl;
			continue;
		}
		std::cout << "read " << read.name << " took " << alignment.elapsedMilliseconds << "ms" << std::endl;
		if (alignment.alignmentFailed || alignment.alignment.score == std::numeric_limits<int32_t>::max())
		{
			std::cout << "read " << read.name << " alignment failed" << std::endl;
			std::cerr << "read " << read.name << " alignment failed" << std::endl;
			continue;
		}
		std::cout << "read " << read.name << " score " << alignment.alignment.score << std::endl;
		if (alignment.alignment.score > read.sequence.size() * 0.25) std::cerr << "read " << read.name << " score is poor: " << alignment.alignment.score << std::endl;
		std::cout << "read " << read.name << " alignment positions: " << alignment.alignmentStart << "-" << alignment.alignmentEnd << " (read " << read.sequence.size() << "bp)" << std::endl;
		// replaceDigraphNodeIdsWithOriginalNodeIds, Aligner.cpp:83-91
		for (auto& m : alignment.alignment.path.mapping) m.position.node_id = (int)m.position.node_id / 2;
		alignments.push_back(alignment.alignment);
		std::cout << "thread " << threadnum << " successfully aligned read " << read.name << " with " << alignment.cellsProcessed << " cells" << std::endl;
		std::string filename = sanitize("alignment_" + std::to_string(threadnum) + "_" + read.name + ".gam");
		std::cout << "write alignment to " << filename << std::endl;
		vgcodec::WriteAlignmentFile(filename, { alignments.back() });
		std::cout << "alignment write finished" << std::endl;
		std::string tracefilename = sanitize("trace_" + std::to_string(threadnum) + "_" + read.name + ".trace");
		std::cout << "write trace to " << tracefilename << std::endl;
		{
			// writeTrace, Aligner.cpp:93-100
			std::ofstream file(tracefilename);
			for (auto& t : alignment.trace) file << t.nodeID << " " << t.offset << " " << (t.reverse ? 1 : 0) << " " << t.readpos << " " << (int)t.type << " " << t.graphChar << " " << t.readChar << std::endl;
		}
		std::cout << "trace write finished" << std::endl;
	}
	std::cout << "thread " << threadnum << " finished with " << alignments.size() << " alignments" << std::endl;
	std::cerr << "final result has " << alignments.size() << " alignments" << std::endl;
	if (alignmentFile != "") vgcodec::WriteAlignmentFile(alignmentFile, alignments);
	if (auggraphFile != "")
	{
		// the reference reads the graph file as a vg graph here whatever its type (Aligner.cpp:317 CommonUtils::LoadVGGraph)
		if (graphFile.size() >= 3 && graphFile.substr(graphFile.size() - 3) == ".vg") writeAugmentedGraph(graphFile, auggraphFile, alignments);
		else std::cerr << "-A needs a .vg graph (the reference parses the graph file as vg for the augmented graph)" << std::endl;
	}
	ReleaseAlignerEngine(graph);
	return 0;
}
