#include "ga_host.h"
#include <algorithm>
#include <chrono>
#include <limits>
#include <stdexcept>

namespace ga
{

static const uint32_t FLAG_STREAM_ERROR = 1;   // a stream hit a hard limit (see ga_stream_out.status)
static const uint32_t FLAG_BAD_SEED = 2;       // seed node not in graph / position outside read
static const uint32_t FLAG_BAD_CHAR = 4;       // read holds a character the reference aborts on
static const uint32_t FLAG_CYCLIC = 8;         // some band held a cyclic component
static const uint32_t FLAG_TIE = 16;           // final-slice minimum tied across nodes

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

static void appendPart(std::vector<uint8_t>& parts, const std::string& part, ga_stream_in& in)
{
	// 'N' padding to a multiple of 64 rows (GraphAligner.h:2989-2998,3006-3016)
	size_t padded = (part.size() + 63) / 64 * 64;
	in.seqOff = parts.size();
	in.partLen = (uint32_t)padded;
	parts.insert(parts.end(), part.begin(), part.end());
	parts.insert(parts.end(), padded - part.size(), (uint8_t)'N');
}

BatchPlan::BatchPlan(const AlignmentGraph& graph, const std::vector<ReadInput>& reads)
{
	firstSeedOfRead.reserve(reads.size() + 1);
	size_t totalBytes = 0;
	for (auto& r : reads) totalBytes += (r.sequence->size() + 128) * std::max<size_t>(1, r.seeds->size());
	parts.reserve(totalBytes);
	for (size_t ri = 0; ri < reads.size(); ri++)
	{
		firstSeedOfRead.push_back((uint32_t)seeds.size());
		const std::string& seq = *reads[ri].sequence;
		bool badChar = false;
		for (char c : seq)
		{
			if (!ValidReadChar(c)) { badChar = true; break; }
		}
		for (size_t si = 0; si < reads[ri].seeds->size(); si++)
		{
			const SeedHit& hit = (*reads[ri].seeds)[si];
			SeedPlan sp;
			sp.read = (uint32_t)ri;
			sp.seed = (uint32_t)si;
			sp.fwStream = -1;
			sp.bwStream = -1;
			sp.invalid = false;
			int nodeId = std::get<0>(hit);
			size_t pos = std::get<1>(hit);
			bool backwards = std::get<2>(hit);
			if (badChar || !graph.HasNode(nodeId * 2) || !graph.HasNode(nodeId * 2 + 1) || pos >= seq.size())
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
				ga_stream_in in;
				in.startNode = (uint32_t)backwardNode;
				appendPart(parts, ReverseComplement(seq.substr(0, pos + graph.DBGOverlap)), in);
				sp.bwStream = (int64_t)streams.size();
				streams.push_back(in);
			}
			if (pos < seq.size() - 1)
			{
				ga_stream_in in;
				in.startNode = (uint32_t)forwardNode;
				appendPart(parts, seq.substr(pos), in);
				sp.fwStream = (int64_t)streams.size();
				streams.push_back(in);
			}
			seeds.push_back(sp);
		}
	}
	firstSeedOfRead.push_back((uint32_t)seeds.size());
}

DirectionTrace DecodeStream(const AlignmentGraph& graph, const ga_stream_in& in, const ga_stream_out& out, const uint32_t* arena)
{
	(void)in;
	DirectionTrace result;
	if (out.status != GA_OK || out.nSlices <= 0) return result;
	result.present = true;
	result.score = out.score;
	result.nSlices = (size_t)out.nSlices;
	const uint32_t* moves = arena + out.traceOff;
	const uint32_t* path = moves + (out.nMoves + 15) / 16;
	std::vector<MatrixPos>& t = result.trace;
	t.reserve(out.nMoves + 1);
	MatrixPos cur;
	cur.node = out.endNode;
	cur.off = out.endOff;
	cur.j = (size_t)out.nSlices * 64 - 1;
	t.push_back(cur);
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
		t.push_back(cur);
	}
	std::reverse(t.begin(), t.end());
	return result;
}

namespace
{

struct PiecewiseTrace
{
	int32_t fwScore = 0, bwScore = 0;
	std::vector<MatrixPos> fw, bw;
	size_t estimatedCorrectlyAligned = 0;
	uint64_t wordColumns = 0;
};

// addAlignmentNodes, GraphAligner.h:594-634
void addTried(std::vector<std::tuple<size_t, size_t, size_t>>& tried, const std::vector<MatrixPos>& trace)
{
	if (trace.empty()) return;
	size_t oldNode = trace[0].node;
	size_t startIndex = trace[0].j, endIndex = trace[0].j;
	for (size_t i = 1; i < trace.size(); i++)
	{
		if (trace[i].node != oldNode)
		{
			tried.emplace_back(startIndex, endIndex, oldNode);
			startIndex = trace[i].j;
			oldNode = trace[i].node;
		}
		endIndex = trace[i].j;
	}
	tried.emplace_back(startIndex, endIndex, oldNode);
}

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

// getTraceInfoInner, GraphAligner.h:718-780
void traceInfoInner(const AlignmentGraph& graph, const std::string& sequence, const std::vector<MatrixPos>& trace, std::vector<AlignmentResult::TraceItem>& result)
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
		item.readChar = sequence[np.j];
		if (np.j == op.j) item.type = AlignmentResult::DELETION;
		else if (sameColumn && !diagonal) item.type = AlignmentResult::INSERTION;
		else item.type = charMatch(sequence[np.j], item.graphChar) ? AlignmentResult::MATCH : AlignmentResult::MISMATCH;
		result.push_back(item);
	}
}

AlignmentResult emptyAlignment()
{
	AlignmentResult r;
	r.alignment.score = std::numeric_limits<int32_t>::max();
	r.alignmentFailed = true;
	return r;
}

// traceToAlignment, GraphAligner.h:782-847 (one Mapping per node run, one Edit per mapping, final mapping's
// from_length without the +1)
AlignmentResult traceToAlignment(const AlignmentGraph& graph, const std::string& seq_id, const std::string& sequence, int32_t score, const std::vector<MatrixPos>& trace)
{
	AlignmentResult r;
	r.alignment.name = seq_id;
	r.alignment.score = score;
	r.alignment.sequence = sequence;
	r.alignmentFailed = true;
	if (trace.empty()) return r;
	size_t pos = 0;
	size_t oldNode = trace[0].node;
	while (oldNode == graph.DummyNodeStart())
	{
		pos++;
		if (pos == trace.size()) return emptyAlignment();
		oldNode = trace[pos].node;
	}
	if (oldNode == graph.DummyNodeEnd()) return emptyAlignment();
	int rank = 0;
	r.alignment.path.mapping.emplace_back();
	{
		vg::Mapping& m = r.alignment.path.mapping.back();
		m.rank = rank;
		m.position.node_id = graph.NodeID(oldNode);
		m.position.is_reverse = graph.Reverse(oldNode);
		m.position.offset = trace[pos].off;
	}
	MatrixPos btNodeStart = trace[pos], btNodeEnd = trace[pos], btBeforeNode = trace[pos];
	for (; pos < trace.size(); pos++)
	{
		if (trace[pos].node == graph.DummyNodeEnd()) break;
		if (trace[pos].node == oldNode)
		{
			btNodeEnd = trace[pos];
			continue;
		}
		vg::Edit e;
		e.from_length = (int32_t)(btNodeEnd.off - btNodeStart.off + 1);
		e.to_length = (int32_t)(btNodeEnd.j - btBeforeNode.j);
		e.sequence = sequence.substr(btNodeStart.j, btNodeEnd.j - btBeforeNode.j);
		e.read_start = btNodeStart.j;
		r.alignment.path.mapping.back().edit.push_back(e);
		oldNode = trace[pos].node;
		btBeforeNode = btNodeEnd;
		btNodeStart = trace[pos];
		btNodeEnd = trace[pos];
		rank++;
		r.alignment.path.mapping.emplace_back();
		vg::Mapping& m = r.alignment.path.mapping.back();
		m.rank = rank;
		m.position.node_id = graph.NodeID(oldNode);
		m.position.is_reverse = graph.Reverse(oldNode);
	}
	vg::Edit e;
	e.from_length = (int32_t)(btNodeEnd.off - btNodeStart.off);
	e.to_length = (int32_t)(btNodeEnd.j - btBeforeNode.j);
	e.sequence = sequence.substr(btNodeStart.j, btNodeEnd.j - btBeforeNode.j);
	e.read_start = btNodeStart.j;
	r.alignment.path.mapping.back().edit.push_back(e);
	r.alignmentFailed = false;
	return r;
}

// mergeAlignments, GraphAligner.h:648-688
AlignmentResult mergeAlignments(const AlignmentGraph& graph, const AlignmentResult& first, const AlignmentResult& second)
{
	if (first.alignmentFailed) return second;
	if (second.alignmentFailed) return first;
	if (first.alignment.path.mapping.empty()) return second;
	if (second.alignment.path.mapping.empty()) return first;
	AlignmentResult fin;
	fin.alignmentFailed = false;
	fin.alignment = first.alignment;
	fin.alignment.score = first.alignment.score + second.alignment.score;
	size_t start = 0;
	const vg::Position& firstEnd = first.alignment.path.mapping.back().position;
	const vg::Position& secondStart = second.alignment.path.mapping.front().position;
	size_t firstEndNode = graph.Lookup((int)firstEnd.node_id);
	size_t secondStartNode = graph.Lookup((int)secondStart.node_id);
	if (firstEnd.node_id == secondStart.node_id && firstEnd.is_reverse == secondStart.is_reverse) start = 1;
	else if (graph.HasOutNeighbor(firstEndNode, secondStartNode)) start = 0;
	// else: the reference only logs "Piecewise alignments can't be merged!" and appends everything
	for (size_t i = start; i < second.alignment.path.mapping.size(); i++) fin.alignment.path.mapping.push_back(second.alignment.path.mapping[i]);
	return fin;
}

}

AlignmentResult AssembleRead(const AlignmentGraph& graph, const ReadInput& read, const BatchPlan& plan, uint32_t readIndex,
	const std::vector<ga_stream_out>& outs, const std::vector<uint32_t>& arena)
{
	const std::string& sequence = *read.sequence;
	uint32_t first = plan.firstSeedOfRead[readIndex], last = plan.firstSeedOfRead[readIndex + 1];
	std::vector<std::tuple<size_t, size_t, size_t>> tried;
	bool hasAlignment = false;
	PiecewiseTrace best;
	size_t bestSeedPos = 0;
	uint32_t flags = 0;
	uint64_t wordColumns = 0;
	for (uint32_t k = first; k < last; k++)
	{
		const BatchPlan::SeedPlan& sp = plan.seeds[k];
		if (sp.fwStream >= 0) wordColumns += outs[sp.fwStream].wordColumns;
		if (sp.bwStream >= 0) wordColumns += outs[sp.bwStream].wordColumns;
	}
	for (uint32_t k = first; k < last; k++)
	{
		const BatchPlan::SeedPlan& sp = plan.seeds[k];
		const SeedHit& hit = (*read.seeds)[sp.seed];
		if (sp.invalid)
		{
			// reference: nodeLookup.at / substr throw std::out_of_range (GraphAligner.h:423), or abort on a bad character
			AlignmentResult r = emptyAlignment();
			bool badChar = false;
			for (char c : sequence) badChar = badChar || !ValidReadChar(c);
			r.flags = badChar ? FLAG_BAD_CHAR : FLAG_BAD_SEED;
			return r;
		}
		size_t nodeIndex = graph.Lookup(std::get<0>(hit) * 2);
		size_t pos = std::get<1>(hit);
		bool already = false;
		for (auto& t : tried)
		{
			if (std::get<0>(t) <= pos && std::get<1>(t) >= pos && std::get<2>(t) == nodeIndex) { already = true; break; }
		}
		if (already) continue;   // "seed i already aligned", GraphAligner.h:425-429
		PiecewiseTrace pw;
		// getPiecewiseTracesFromSplit, GraphAligner.h:3039-3098
		size_t splitIndex = pos;
		bool streamError = false;
		size_t fwSlices = 0, bwSlices = 0;
		if (sp.fwStream >= 0)
		{
			const ga_stream_out& o = outs[sp.fwStream];
			if (o.status != GA_OK && o.status != GA_EMPTY) streamError = true;
			if (o.cyclicSlices) flags |= FLAG_CYCLIC;
			DirectionTrace d = DecodeStream(graph, plan.streams[sp.fwStream], o, arena.data());
			if (d.present)
			{
				fwSlices = d.nSlices;
				pw.fwScore = d.score;
				pw.fw.swap(d.trace);
				size_t backtraceableSize = sequence.size() - splitIndex - graph.DBGOverlap;
				while (!pw.fw.empty() && pw.fw.back().j >= backtraceableSize) pw.fw.pop_back();
			}
		}
		if (sp.bwStream >= 0)
		{
			const ga_stream_out& o = outs[sp.bwStream];
			if (o.status != GA_OK && o.status != GA_EMPTY) streamError = true;
			if (o.cyclicSlices) flags |= FLAG_CYCLIC;
			DirectionTrace d = DecodeStream(graph, plan.streams[sp.bwStream], o, arena.data());
			if (d.present)
			{
				bwSlices = d.nSlices;
				pw.bwScore = d.score;
				pw.bw.swap(d.trace);
				while (!pw.bw.empty() && pw.bw.back().j >= splitIndex) pw.bw.pop_back();
				// reverseTrace, GraphAligner.h:3026-3037
				std::reverse(pw.bw.begin(), pw.bw.end());
				for (auto& p : pw.bw)
				{
					size_t other = graph.GetReverseNode(p.node);
					p.off = (uint32_t)(graph.NodeLength(other) - 1 - p.off);
					p.node = (uint32_t)other;
					p.j = (splitIndex - 1) - p.j;
				}
				// the forward rows are shifted only inside this branch in the reference (GraphAligner.h:3090-3093)
				for (auto& p : pw.fw) p.j += splitIndex;
			}
		}
		if (streamError)
		{
			AlignmentResult r = emptyAlignment();
			r.flags = flags | FLAG_STREAM_ERROR;
			r.wordColumns = wordColumns;
			return r;
		}
		pw.estimatedCorrectlyAligned = (fwSlices + bwSlices) * 64;
		addTried(tried, pw.fw);
		addTried(tried, pw.bw);
		if (!hasAlignment || pw.estimatedCorrectlyAligned > best.estimatedCorrectlyAligned)
		{
			best = std::move(pw);
			bestSeedPos = pos;
			hasAlignment = true;
		}
	}
	if (!hasAlignment)
	{
		AlignmentResult r = emptyAlignment();
		r.flags = flags;
		r.wordColumns = wordColumns;
		return r;
	}
	// getTraceInfo, GraphAligner.h:690-716
	std::vector<AlignmentResult::TraceItem> traceVector;
	if (!best.bw.empty()) traceInfoInner(graph, sequence, best.bw, traceVector);
	if (!best.bw.empty() && !best.fw.empty())
	{
		const MatrixPos& p = best.fw[0];
		AlignmentResult::TraceItem item;
		item.type = AlignmentResult::FORWARDBACKWARDSPLIT;
		item.nodeID = graph.NodeID(p.node) / 2;
		item.reverse = p.node % 2 == 1;   // node INDEX parity, as the reference writes it (GraphAligner.h:704)
		item.offset = p.off;
		item.readpos = p.j;
		item.graphChar = graph.NodeSequences(graph.NodeStart(p.node) + p.off);
		item.readChar = sequence[p.j];
		traceVector.push_back(item);
	}
	if (!best.fw.empty()) traceInfoInner(graph, sequence, best.fw, traceVector);

	AlignmentResult fwresult = traceToAlignment(graph, *read.name, sequence, best.fwScore, best.fw);
	AlignmentResult bwresult = traceToAlignment(graph, *read.name, sequence, best.bwScore, best.bw);
	if (fwresult.alignmentFailed && bwresult.alignmentFailed)
	{
		AlignmentResult r = emptyAlignment();
		r.flags = flags;
		r.wordColumns = wordColumns;
		return r;
	}
	AlignmentResult result = mergeAlignments(graph, bwresult, fwresult);
	result.trace.swap(traceVector);
	size_t lastAligned = !best.bw.empty() ? best.bw[0].j : bestSeedPos;
	result.alignment.query_position = (int32_t)lastAligned;
	result.alignmentStart = lastAligned;
	result.alignmentEnd = lastAligned + best.estimatedCorrectlyAligned;
	result.flags = flags;
	result.wordColumns = wordColumns;
	return result;
}

std::vector<AlignmentResult> AlignBatch(DeviceCtx* ctx, const AlignmentGraph& graph, const std::vector<ReadInput>& reads, int initialBandwidth, int rampBandwidth, BatchStats* stats)
{
	if (!graph.Finalized()) throw std::logic_error("AlignBatch: graph not finalized");
	BatchPlan plan(graph, reads);
	std::vector<ga_stream_out> outs;
	std::vector<uint32_t> arena;
	ExecuteStreams(ctx, plan.streams, plan.parts, initialBandwidth, rampBandwidth, outs, arena, stats);
	std::vector<AlignmentResult> results(reads.size());
	for (size_t i = 0; i < reads.size(); i++)
	{
		if (reads[i].seeds->empty())
		{
			results[i] = emptyAlignment();   // Aligner.cpp:131-138 "has no seed hits"
			continue;
		}
		results[i] = AssembleRead(graph, reads[i], plan, (uint32_t)i, outs, arena);
	}
	if (stats)
	{
		stats->streams += plan.streams.size();
		for (auto& o : outs) stats->wordColumns += o.wordColumns;
	}
	return results;
}

}
