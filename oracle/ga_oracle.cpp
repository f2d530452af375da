// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or executed from the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may run this program.
//
// ga_oracle: a plain, cell-by-cell CPU restatement of the reference's seeded alignment path
// (GraphAlignerWrapper.h:54 -> GraphAligner.h).  Nothing here is bit-parallel: every DP cell is an
// int, every rule is written out, each function cites the reference lines it restates.  It is pinned
// against the reference itself: tests/test_cpu.py checks it against tests/golden/*.expected, which are
// outputs of the UNMODIFIED reference sources (oracle/_ref/ref_align, see tests/golden/make_golden.py).
//
// Reads the same .gacase files and prints the same lines as oracle/ref_driver.cpp.
//
// Known limits (stated, not hidden): bands >= 200 000 bp (the reference's calculateSliceAlternate /
// BacktraceOverride, GraphAligner.h:2148-2329,167-354) are not restated.  The -B ramp redo (GraphAligner.h:2648-2719) IS, with
// the sqrt checkpoints and the traceback through re-computed slices (getSlicesFromTable, :2858-2943) as the reference has
// them - including the pending checkpoint a redo does not rewind (fixtures ramp_redo, ramp_stale, ramp_stale_long).  Inside a cyclic band component the reference's per-node minimum and tie order depend on
// its work-list schedule (confirmedRows, GraphAligner.h:1355-1416,2364-2420), here the fix point is used - the two
// fixtures that pin that schedule (tests/golden/cyclic_*) are checked against oracle/_ref/ref_align only.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <queue>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

static const int INF = 1000000000;

// ---------------------------------------------------------------------------------------------------------------
// graph: AlignmentGraph.cpp:11-138 (dummy nodes, AddNode, AddEdgeNodeId with duplicate suppression, Finalize)
// ---------------------------------------------------------------------------------------------------------------
struct Graph
{
	std::vector<size_t> nodeStart;
	std::vector<int> nodeIDs;
	std::vector<bool> reverse;
	std::string seq;   // one char per bp, '-' for the dummies
	std::vector<std::vector<size_t>> in, out;
	std::unordered_map<int, size_t> lookup;
	int overlap = 0;
	Graph() { addRaw(0, "-", false); }
	void addRaw(int id, const std::string& s, bool rev)
	{
		nodeIDs.push_back(id);
		nodeStart.push_back(seq.size());
		reverse.push_back(rev);
		in.emplace_back();
		out.emplace_back();
		seq += s;
	}
	void addNode(int id, const std::string& s, bool rev)
	{
		if (lookup.count(id)) return;   // AlignmentGraph.cpp:49-51
		lookup[id] = nodeStart.size();
		addRaw(id, s, rev);
	}
	void addEdge(int from, int to)
	{
		size_t f = lookup.at(from), t = lookup.at(to);
		if (std::find(in[t].begin(), in[t].end(), f) == in[t].end()) in[t].push_back(f);      // AlignmentGraph.cpp:104
		if (std::find(out[f].begin(), out[f].end(), t) == out[f].end()) out[f].push_back(t);  // AlignmentGraph.cpp:105
	}
	void finalize() { addRaw(0, "-", false); }   // AlignmentGraph.cpp:108-118
	size_t size() const { return nodeStart.size(); }
	size_t start(size_t n) const { return nodeStart[n]; }
	size_t end(size_t n) const { return n + 1 == nodeStart.size() ? seq.size() : nodeStart[n + 1]; }
	size_t len(size_t n) const { return end(n) - start(n); }
	size_t indexToNode(size_t w) const   // AlignmentGraph.cpp:226-234
	{
		return (size_t)(std::upper_bound(nodeStart.begin(), nodeStart.end(), w) - nodeStart.begin()) - 1;
	}
	size_t reverseNode(size_t n) const   // AlignmentGraph.cpp:199-214
	{
		int id = nodeIDs[n];
		return lookup.at(id % 2 == 1 ? (id / 2) * 2 : (id / 2) * 2 + 1);
	}
	size_t reversePosition(size_t w) const   // AlignmentGraph.cpp:216-224
	{
		size_t n = indexToNode(w), o = reverseNode(n);
		return (end(o) - 1) - (w - start(n));
	}
};

// CommonUtils.cpp:60-136 (the NDEBUG behaviour: an unknown character is dropped)
static std::string reverseComplement(const std::string& s)
{
	static const std::map<char, char> comp = { {'A','T'},{'C','G'},{'T','A'},{'G','C'},{'N','N'},{'U','A'},{'R','Y'},{'Y','R'},{'K','M'},{'M','K'},{'S','S'},{'W','W'},{'B','V'},{'V','B'},{'D','H'},{'H','D'} };
	std::string r;
	for (size_t i = s.size(); i-- > 0;)
	{
		auto f = comp.find((char)toupper(s[i]));
		if (f != comp.end()) r += f->second;
	}
	return r;
}

// GraphAligner.h:2039-2110
static bool characterMatch(char readChar, char graphChar)
{
	static const std::map<char, std::string> iupac = { {'A',"A"},{'C',"C"},{'G',"G"},{'T',"T"},{'N',"ACGT"},{'R',"AG"},{'Y',"CT"},{'K',"GT"},{'M',"CA"},{'S',"CG"},{'W',"AT"},{'B',"CGT"},{'D',"AGT"},{'H',"ACT"},{'V',"ACG"} };
	auto f = iupac.find((char)toupper(readChar));
	if (f == iupac.end()) { fprintf(stderr, "ga_oracle: illegal read character %c (the reference aborts)\n", readChar); exit(3); }
	return f->second.find(graphChar) != std::string::npos;
}

// ---------------------------------------------------------------------------------------------------------------
// one 64-row slice of the DP, cell by cell
// ---------------------------------------------------------------------------------------------------------------
struct Column
{
	int sbs = 0;                 // "scoreBeforeStart": value in row -1 (= row 63 of the previous slice, recomputed)
	int v[64];                   // cell values of rows 0..63
	bool scoreBeforeExists = false;
	int scoreEnd() const { return v[63]; }
};

// AlignmentCorrectnessEstimation.cpp:6-30,71-89 - same constants, same association of the sums
struct Hmm
{
	double correct = log(0.8), wrong = log(0.2);
	bool correctFromCorrect = false, falseFromCorrect = false;
	bool currentlyCorrect() const { return correct > wrong; }
	Hmm next(int mismatches) const
	{
		static const double cm = log(0.2), cM = log(1.0 - 0.2), fm = log(0.5), fM = log(1.0 - 0.5);
		static const double f2c = log(0.00001), f2f = log(1.0 - 0.00001), c2f = log(0.000000000000001), c2c = log(1.0 - 0.000000000000001);
		static std::vector<double> lf;
		if (lf.empty()) { lf.push_back(0); for (int i = 1; i <= 64; i++) lf.push_back(lf.back() + log(i)); }
		Hmm r;
		r.correctFromCorrect = correct + c2c >= wrong + f2c;
		r.falseFromCorrect = correct + c2f >= wrong + f2f;
		double nc = std::max(correct + c2c, wrong + f2c);
		double nf = std::max(correct + c2f, wrong + f2f);
		double choose = lf[64] - lf[mismatches] - lf[64 - mismatches];
		nc += choose + mismatches * cm + (64 - mismatches) * cM;
		nf += choose + mismatches * fm + (64 - mismatches) * fM;
		r.correct = nc;
		r.wrong = nf;
		return r;
	}
};

struct Slice
{
	long j = -64;                                   // first row
	std::vector<size_t> nodes;                      // band, in the reference's band order
	std::map<size_t, std::vector<Column>> cols;     // node -> its columns
	std::map<size_t, int> nodeMin;
	int minScore = 0;
	std::vector<size_t> minScoreIndex;              // tied minimum columns, in evaluation order
	Hmm hmm;                                        // DPSlice::correctness after this slice
	bool has(size_t n) const { return cols.count(n) != 0; }
};

struct Aligner
{
	const Graph& g;
	int b, B;
	Aligner(const Graph& g, int b, int B) : g(g), b(b), B(B) {}

	// projectForwardFromMinScore, GraphAligner.h:1110-1159.  The previous slice's node map is a real
	// std::unordered_map filled in band order (NodeSlice.h:730-733), walked in its own order (GraphAligner.h:1117).
	std::vector<size_t> selectBand(const Slice& prev, int bandwidth) const
	{
		std::unordered_map<size_t, int> walk;
		for (size_t n : prev.nodes) walk[n] = 0;
		const int expand = bandwidth + 64;
		std::unordered_map<size_t, size_t> distances;
		std::vector<size_t> result;
		typedef std::pair<int, size_t> Item;   // (priority, node); only the priority is compared, like NodeWithPriority
		auto cmp = [](const Item& a, const Item& c) { return a.first > c.first; };
		std::priority_queue<Item, std::vector<Item>, decltype(cmp)> queue(cmp);
		size_t width = 0;
		for (auto& pair : walk)
		{
			size_t node = pair.first;
			if (prev.nodeMin.at(node) > prev.minScore + bandwidth) continue;
			distances[node] = 0;
			result.push_back(node);
			width += g.len(node);
			if (width >= 200000) { fprintf(stderr, "ga_oracle: band >= 200000 bp needs the alternate method (not restated)\n"); exit(4); }
			int endscore = prev.cols.at(node).back().scoreEnd();
			if (endscore > prev.minScore + expand) continue;
			for (size_t nb : g.out[node]) queue.emplace(endscore - prev.minScore + 1, nb);
		}
		while (!queue.empty())
		{
			Item top = queue.top();
			if (top.first > expand) break;
			queue.pop();
			if (distances.count(top.second) && (int)distances[top.second] <= top.first) continue;
			width += g.len(top.second);
			distances[top.second] = top.first;
			result.push_back(top.second);
			if (width >= 200000) { fprintf(stderr, "ga_oracle: band >= 200000 bp needs the alternate method (not restated)\n"); exit(4); }
			for (size_t nb : g.out[top.second]) queue.emplace(top.first + (int)g.len(top.second), nb);
		}
		return result;
	}

	// getStronglyConnectedComponents, GraphAligner.h:1759-1856 (Tarjan, band order, outNeighbors order)
	std::vector<std::vector<size_t>> components(const Slice& cur) const
	{
		std::vector<std::vector<size_t>> result;
		std::map<size_t, size_t> index, low;
		std::map<size_t, bool> onStack;
		std::vector<size_t> stack;
		size_t counter = 0;
		struct Frame { size_t node; size_t next; };
		for (size_t root : cur.nodes)
		{
			if (index.count(root)) continue;
			std::vector<Frame> call;
			call.push_back({ root, 0 });
			index[root] = low[root] = counter++;
			stack.push_back(root);
			onStack[root] = true;
			while (!call.empty())
			{
				Frame& f = call.back();
				if (f.next < g.out[f.node].size())
				{
					size_t nb = g.out[f.node][f.next++];
					if (!cur.has(nb)) continue;
					if (!index.count(nb))
					{
						index[nb] = low[nb] = counter++;
						stack.push_back(nb);
						onStack[nb] = true;
						call.push_back({ nb, 0 });
					}
					else if (onStack[nb]) low[f.node] = std::min(low[f.node], index[nb]);
					continue;
				}
				size_t node = f.node;
				call.pop_back();
				if (!call.empty()) low[call.back().node] = std::min(low[call.back().node], low[node]);
				if (low[node] == index[node])
				{
					result.emplace_back();
					size_t back;
					do
					{
						back = stack.back();
						stack.pop_back();
						onStack[back] = false;
						result.back().push_back(back);
					} while (back != node);
				}
			}
		}
		return result;
	}

	// One candidate for a column from one left neighbour column: the word step getNextSlice (GraphAligner.h:1349-1399)
	// written per cell.  left == nullptr means the neighbour is only in the previous band: a synthetic column that
	// climbs by one per row from its previous end score, which may match only in row 0 (GraphAligner.h:1294-1301).
	void candidate(Column& out, const Column* left, int syntheticScore, bool leftExists, bool upleftInside, bool diagInside, bool previousEq,
		int upleftRow62, char graphChar, const std::string& part, long j) const
	{
		int leftSbs = left ? left->sbs : syntheticScore;
		int sbs = leftSbs + 1;
		if (upleftInside) sbs = std::min(sbs, upleftRow62 + (previousEq ? 0 : 1));   // GraphAligner.h:1361-1370
		out.sbs = sbs;
		int up = sbs, leftPrevRow = leftSbs;
		for (int r = 0; r < 64; r++)
		{
			int leftHere = left ? left->v[r] : syntheticScore + r + 1;
			bool match = characterMatch(part[j + r], graphChar);
			if (r == 0 && (!leftExists || !diagInside)) match = false;   // Eq &= ~1, GraphAligner.h:1358-1360
			if (!left && r > 0) match = false;                            // EqHere &= 1, GraphAligner.h:1301
			int val = std::min(std::min(up + 1, leftHere + 1), leftPrevRow + (match ? 0 : 1));
			out.v[r] = val;
			up = val;
			leftPrevRow = leftHere;
		}
	}

	static void mergeMin(Column& a, const Column& c)   // WordSlice::mergeTwoSlices, WordSlice.h:361-421: cell-wise minimum
	{
		a.sbs = std::min(a.sbs, c.sbs);
		for (int r = 0; r < 64; r++) a.v[r] = std::min(a.v[r], c.v[r]);
	}

	// row -1 scores of one component by shortest paths, then reset (forceComponentZeroRow, GraphAligner.h:1903-1995)
	void forceZeroRow(Slice& cur, const Slice& prev, const std::vector<size_t>& comp, const std::map<size_t, size_t>& compOf, size_t compIndex) const
	{
		typedef std::pair<int, size_t> Item;
		std::priority_queue<Item, std::vector<Item>, std::greater<Item>> queue;
		for (size_t node : comp)
		{
			auto& c = cur.cols[node];
			for (auto& col : c) col.sbs = INF;
			if (prev.has(node)) c[0].sbs = prev.cols.at(node)[0].scoreEnd();
			for (size_t nb : g.in[node])
			{
				if (!cur.has(nb) && !prev.has(nb)) continue;
				if (compOf.count(nb) && compOf.at(nb) == compIndex) continue;
				if (cur.has(nb)) c[0].sbs = std::min(c[0].sbs, cur.cols[nb].back().sbs + 1);
				if (prev.has(nb)) c[0].sbs = std::min(c[0].sbs, prev.cols.at(nb).back().scoreEnd() + 1);
			}
			if (c[0].sbs == INF) continue;
			for (size_t i = 1; i < c.size(); i++)
			{
				c[i].sbs = c[i - 1].sbs + 1;
				if (prev.has(node)) c[i].sbs = std::min(c[i].sbs, prev.cols.at(node)[i].scoreEnd());
			}
			for (size_t nb : g.out[node])
			{
				if (compOf.count(nb) && compOf.at(nb) == compIndex) queue.emplace(c.back().sbs + 1, nb);
			}
		}
		while (!queue.empty())
		{
			Item top = queue.top();
			queue.pop();
			auto& c = cur.cols[top.second];
			int score = top.first;
			bool endUpdated = true;
			for (auto& col : c)
			{
				if (col.sbs <= score) { endUpdated = false; break; }
				col.sbs = score++;
			}
			if (!endUpdated) continue;
			for (size_t nb : g.out[top.second])
			{
				if (compOf.count(nb) && compOf.at(nb) == compIndex) queue.emplace(score, nb);
			}
		}
		for (size_t node : comp)
		{
			auto& c = cur.cols[node];
			for (size_t i = 0; i < c.size(); i++)
			{
				for (int r = 0; r < 64; r++) c[i].v[r] = c[i].sbs + r + 1;   // VP = all ones
				c[i].scoreBeforeExists = prev.has(node) && prev.cols.at(node)[i].scoreEnd() == c[i].sbs;   // scoreEndExists is always true on this path
			}
		}
	}

	// calculateNode, GraphAligner.h:1457-1573 (+ getNodeStartSlice :1270-1315, sources :1317-1347).  Returns whether a cell changed.
	bool calcNode(Slice& cur, const Slice& prev, size_t node, const std::string& part) const
	{
		std::vector<Column>& c = cur.cols[node];
		std::vector<Column> before = c;
		const bool inPrev = prev.has(node);
		const long j = cur.j;
		for (size_t k = 0; k < c.size(); k++)
		{
			char graphChar = g.seq[g.start(node) + k];
			bool previousEq = (j == 0 && inPrev) || (j > 0 && graphChar == part[j - 1]);   // exact compare, GraphAligner.h:1503,1540
			Column res;
			res.scoreBeforeExists = c[k].scoreBeforeExists;
			bool have = false;
			if (k == 0)
			{
				for (size_t nb : g.in[node])
				{
					if (!cur.has(nb) && !prev.has(nb)) continue;
					bool foundOneUp = prev.has(nb);
					int upRow62 = foundOneUp ? prev.cols.at(nb).back().v[62] : 0;
					Column cand;
					if (cur.has(nb)) candidate(cand, &cur.cols[nb].back(), 0, cur.cols[nb].back().scoreBeforeExists, c[0].scoreBeforeExists && foundOneUp, foundOneUp, previousEq, upRow62, graphChar, part, j);
					else candidate(cand, nullptr, prev.cols.at(nb).back().scoreEnd(), true, c[0].scoreBeforeExists && foundOneUp, foundOneUp, previousEq, upRow62, graphChar, part, j);
					if (!have) { res.sbs = cand.sbs; memcpy(res.v, cand.v, sizeof(res.v)); have = true; }
					else mergeMin(res, cand);
				}
				if (!have)
				{
					// source node, GraphAligner.h:1317-1347,1475-1488
					if (!inPrev) { fprintf(stderr, "ga_oracle: band node without predecessor\n"); exit(5); }
					int ps = prev.cols.at(node)[0].scoreEnd();
					res.sbs = ps;
					int first = (j == 0) ? (characterMatch(part[0], graphChar) ? 0 : 1) : 1;
					res.v[0] = ps + first;
					for (int r = 1; r < 64; r++) res.v[r] = res.v[r - 1] + 1;
					res.scoreBeforeExists = true;
					c[0] = res;
					continue;
				}
			}
			else
			{
				int upRow62 = inPrev ? prev.cols.at(node)[k - 1].v[62] : 0;
				candidate(res, &c[k - 1], 0, c[k - 1].scoreBeforeExists, c[k].scoreBeforeExists, c[k - 1].scoreBeforeExists, previousEq, upRow62, graphChar, part, j);
			}
			if (inPrev && res.sbs > prev.cols.at(node)[k].scoreEnd())
			{
				// merge with the vertical ramp from the previous slice, GraphAligner.h:1504-1509,1541-1546
				int top = prev.cols.at(node)[k].scoreEnd();
				res.sbs = top;
				for (int r = 0; r < 64; r++) res.v[r] = std::min(res.v[r], top + r + 1);
				res.scoreBeforeExists = true;
			}
			c[k] = res;
		}
		for (size_t k = 0; k < c.size(); k++)
		{
			if (memcmp(c[k].v, before[k].v, sizeof(c[k].v)) != 0) return true;
		}
		return false;
	}

	// calculateSlice, GraphAligner.h:2331-2451: components in reverse Tarjan order; cyclic ones to their fix point
	void fillSlice(Slice& cur, const Slice& prev, const std::string& part) const
	{
		for (size_t n : cur.nodes) cur.cols[n].assign(g.len(n), Column());
		auto comps = components(cur);
		std::map<size_t, size_t> compOf;
		for (size_t i = 0; i < comps.size(); i++) for (size_t n : comps[i]) compOf[n] = i;
		cur.minScore = INF;
		for (size_t ci = comps.size(); ci-- > 0;)
		{
			forceZeroRow(cur, prev, comps[ci], compOf, ci);
			bool cyclic = comps[ci].size() > 1 || std::find(g.out[comps[ci][0]].begin(), g.out[comps[ci][0]].end(), comps[ci][0]) != g.out[comps[ci][0]].end();
			// UniqueQueue is a LIFO filled in component order (GraphAligner.h:2363, UniqueQueue.h:24-33): the root is evaluated first
			std::vector<size_t> order(comps[ci].rbegin(), comps[ci].rend());
			bool changed = true;
			int guard = 0;
			while (changed)
			{
				changed = false;
				for (size_t n : order) changed = calcNode(cur, prev, n, part) || changed;
				if (!cyclic) break;
				if (++guard > 100000) { fprintf(stderr, "ga_oracle: cyclic component did not converge\n"); exit(6); }
			}
			for (size_t n : order)
			{
				int nodeMin = INF;
				for (auto& col : cur.cols[n]) nodeMin = std::min(nodeMin, col.scoreEnd());
				cur.nodeMin[n] = nodeMin;
				if (nodeMin < cur.minScore) { cur.minScore = nodeMin; cur.minScoreIndex.clear(); }
				if (nodeMin == cur.minScore)
				{
					for (size_t k = 0; k < cur.cols[n].size(); k++)
					{
						if (cur.cols[n][k].scoreEnd() == nodeMin) cur.minScoreIndex.push_back(g.start(n) + k);
					}
				}
			}
		}
	}

	// DPTable, GraphAligner.h:356-368.  `slices` are the sqrt checkpoints (slices[0] = the initial slice, j = -64); the reference keeps
	// only their end scores (getFrozenSqrtEndScores), here a checkpoint is the whole slice - fillSlice reads its last row only.
	struct Table
	{
		std::vector<Slice> slices;
		std::vector<int> bandwidthPerSlice;
		std::vector<Hmm> correctness;   // one per retained slice
		size_t samplingFrequency = 1;
	};

	// DPSlice::EstimatedMemoryUsage, GraphAligner.h:136-139: sizeof(TinySlice) = 4 per cell of the band, 28 per node; the initial
	// slice never has its cells counted (getInitialSliceOnlyOneNode leaves numCells 0)
	size_t memoryUse(const Slice& sl) const
	{
		size_t cells = 0;
		if (sl.j >= 0) for (size_t n : sl.nodes) cells += g.len(n);
		return cells * 4 + sl.nodes.size() * 28;
	}

	// pickMethodAndExtendFill, GraphAligner.h:2473-2521 (bit-vector branch) with the HMM step of fillDPSlice
	Slice nextSlice(const Slice& prev, const std::string& part, int bandwidth) const
	{
		Slice cur;
		cur.j = prev.j + 64;
		cur.nodes = selectBand(prev, bandwidth);
		fillSlice(cur, prev, part);
		cur.hmm = prev.hmm.next(cur.minScore - prev.minScore);
		return cur;
	}

	// getSqrtSlices + removeWronglyAlignedEnd, GraphAligner.h:2554-2856, as written: the -B ramp redo (:2648-2719) swaps back to the
	// remembered slice and pops table.slices, but the pending checkpoint `storeSlice` is NOT rewound - a slice of the abandoned
	// pass can be pushed at the next sampling point (:2772-2786)
	Table forward(const std::string& part, size_t startNode) const
	{
		Table t;
		Slice init;   // getInitialSliceOnlyOneNode, GraphAligner.h:2945-2960
		init.nodes.push_back(startNode);
		init.cols[startNode].assign(g.len(startNode), Column());
		for (auto& col : init.cols[startNode]) { col.sbs = 0; for (int r = 0; r < 64; r++) col.v[r] = 0; }
		init.nodeMin[startNode] = 0;
		init.minScore = 0;
		init.minScoreIndex.push_back(g.end(startNode) - 1);
		const size_t numSlices = part.size() / 64;
		t.samplingFrequency = (size_t)(int)sqrt((double)numSlices);   // getSamplingFrequency, GraphAligner.h:2962-2967
		Slice lastSlice = init, storeSlice = init, rampSlice = init;
		size_t rampRedoIndex = (size_t)-1, rampUntil = 0;
		for (size_t slice = 0; slice < numSlices; slice++)
		{
			const int bandwidth = (rampUntil >= slice) ? B : b;   // GraphAligner.h:2612
			Slice newSlice = nextSlice(lastSlice, part, bandwidth);
			if (rampUntil == slice - 1 || (rampUntil < slice && newSlice.hmm.currentlyCorrect() && newSlice.hmm.falseFromCorrect))   // :2630-2634
			{
				rampSlice = lastSlice;
				rampRedoIndex = slice - 1;
			}
			if (!newSlice.hmm.correctFromCorrect) break;   // :2640-2647
			if (!newSlice.hmm.currentlyCorrect() && rampUntil < slice && B > b)   // :2648-2719
			{
				rampUntil = slice;
				std::swap(slice, rampRedoIndex);
				std::swap(lastSlice, rampSlice);
				while (t.bandwidthPerSlice.size() > slice + 1) t.bandwidthPerSlice.pop_back();
				while (t.correctness.size() > slice + 1) t.correctness.pop_back();
				while (t.slices.size() > 1 && t.slices.back().j > (long)(slice * 64)) t.slices.pop_back();
				continue;
			}
			t.bandwidthPerSlice.push_back(bandwidth);
			t.correctness.push_back(newSlice.hmm);
			if (slice % t.samplingFrequency == 0)   // :2772-2782
			{
				if (t.slices.empty() || storeSlice.j != t.slices.back().j)
				{
					t.slices.push_back(storeSlice);
					storeSlice = newSlice;
				}
			}
			if (memoryUse(newSlice) < memoryUse(storeSlice)) storeSlice = newSlice;   // :2783-2786
			lastSlice = std::move(newSlice);
		}
		// removeWronglyAlignedEnd, GraphAligner.h:2554-2569
		if (!t.correctness.empty())
		{
			bool currentlyCorrect = t.correctness.back().currentlyCorrect();
			while (!currentlyCorrect)
			{
				t.correctness.pop_back();
				t.bandwidthPerSlice.pop_back();
				if (t.correctness.empty()) break;
				currentlyCorrect = t.correctness.back().falseFromCorrect;
			}
			if (t.correctness.empty()) t.slices.clear();
			while (t.slices.size() > 1 && t.slices.back().j >= (long)(t.correctness.size() * 64)) t.slices.pop_back();
		}
		return t;
	}

	// getSlicesFromTable, GraphAligner.h:2858-2943: the slices behind checkpoint startIndex, up to the next checkpoint (or the end),
	// computed again from the checkpoint with the bandwidths the forward pass recorded
	std::vector<Slice> slicesFromTable(const Table& t, const std::string& part, size_t startIndex) const
	{
		const size_t startSlice = (size_t)((t.slices[startIndex].j + 64) / 64);
		size_t endSlice = startIndex == t.slices.size() - 1 ? t.bandwidthPerSlice.size() : (size_t)((t.slices[startIndex + 1].j + 64) / 64);
		if (endSlice <= startSlice || endSlice > t.bandwidthPerSlice.size())
		{
			fprintf(stderr, "ga_oracle: sqrt checkpoints out of order after a ramp redo (the reference reads an empty stretch here)\n");
			exit(8);
		}
		std::vector<Slice> result;
		Slice lastSlice = t.slices[startIndex];
		for (size_t slice = startSlice; slice < endSlice; slice++)
		{
			result.push_back(nextSlice(lastSlice, part, t.bandwidthPerSlice[slice]));
			lastSlice = result.back();
		}
		return result;
	}

	int valueOrMax(const Slice& s, size_t w, int row, int maxv) const   // getValueOrMax, GraphAligner.h:2008-2017
	{
		size_t n = g.indexToNode(w);
		if (!s.has(n)) return maxv;
		return s.cols.at(n)[w - g.start(n)].v[row];
	}

	typedef std::pair<size_t, long> Pos;   // (graph position, read row)

	// pickBacktracePredecessor, GraphAligner.h:493-591
	Pos predecessor(const Slice& slice, const Slice& previous, const std::string& part, Pos pos) const
	{
		int row = (int)(pos.second - slice.j);
		int maxv = (int)part.size();
		size_t node = g.indexToNode(pos.first);
		if (!slice.has(node))
		{
			fprintf(stderr, "ga_oracle: the trace stands on a node the slice does not hold (the reference reads a missing slice here)\n");
			exit(8);
		}
		int here = slice.cols.at(node)[pos.first - g.start(node)].v[row];
		if (pos.second == 0 && previous.has(node) && (here == 0 || here == 1)) return { pos.first, -1 };
		bool match = characterMatch(part[pos.second], g.seq[pos.first]);
		std::vector<size_t> lefts;
		if (pos.first == g.start(node)) { for (size_t nb : g.in[node]) lefts.push_back(g.end(nb) - 1); }
		else lefts.push_back(pos.first - 1);
		for (size_t u : lefts)
		{
			if (valueOrMax(slice, u, row, maxv) == here - 1) return { u, pos.second };
			int diag = row == 0 ? valueOrMax(previous, u, 63, maxv) : valueOrMax(slice, u, row - 1, maxv);
			if (diag == (match ? here : here - 1)) return { u, pos.second - 1 };
		}
		int up = row == 0 ? valueOrMax(previous, pos.first, 63, maxv) : valueOrMax(slice, pos.first, row - 1, maxv);
		if (up == here - 1) return { pos.first, pos.second - 1 };
		fprintf(stderr, "ga_oracle: no backtrace predecessor (the reference aborts here)\n");
		exit(7);
	}

	// getTraceFromTable, GraphAligner.h:894-957, with getTraceFromTableInner / getTraceFromSlice / getSliceBoundaryTrace (:960-1021):
	// from the last checkpoint to the first, the stretch behind each one is computed again and walked; the step across a
	// checkpoint looks at the checkpoint's own slice
	std::pair<int, std::vector<Pos>> trace(const Table& t, const std::string& part) const
	{
		if (t.slices.empty() || t.bandwidthPerSlice.empty()) return { std::numeric_limits<int>::max(), {} };
		int score = 0;
		std::vector<Pos> result;
		for (size_t i = t.slices.size() - 1; i < t.slices.size(); i--)
		{
			if ((size_t)((t.slices[i].j + 64) / 64) == t.bandwidthPerSlice.size())
			{
				score = t.slices.back().minScore;
				result.emplace_back(t.slices.back().minScoreIndex.back(), t.slices.back().j + 63);
				continue;
			}
			const std::vector<Slice> partTable = slicesFromTable(t, part, i);
			if (i == t.slices.size() - 1)
			{
				score = partTable.back().minScore;
				result.emplace_back(partTable.back().minScoreIndex.back(), partTable.back().j + 63);
			}
			for (size_t k = partTable.size() - 1; k < partTable.size(); k--)
			{
				// getTraceFromSlice: down to the slice's first row
				while (result.back().second != partTable[k].j) result.push_back(predecessor(partTable[k], partTable[k], part, result.back()));
				// getSliceBoundaryTrace: along the first row and over the border, into the slice before (the checkpoint for k = 0)
				const Slice& before = k > 0 ? partTable[k - 1] : t.slices[i];
				while (result.back().second == partTable[k].j) result.push_back(predecessor(partTable[k], before, part, result.back()));
			}
		}
		result.pop_back();   // the position in row -1
		std::reverse(result.begin(), result.end());
		return { score, result };
	}
};

// ---------------------------------------------------------------------------------------------------------------
// result assembly: GraphAligner.h:408-491 (seed loop), :594-688, :690-847, :2969-3098
// ---------------------------------------------------------------------------------------------------------------
struct MappingOut { long node_id; int is_reverse; long offset; int from_length; int to_length; };
struct TraceItem { int nodeID; size_t offset; bool reverse; size_t readpos; int type; };
struct ReadResult
{
	bool failed = true;
	int score = 0;
	size_t start = 0, end = 0;
	int qpos = 0;
	std::vector<MappingOut> mappings;
	std::vector<TraceItem> trace;
};

typedef Aligner::Pos Pos;

// traceToAlignment, GraphAligner.h:782-847
static bool traceToMappings(const Graph& g, const std::vector<Pos>& trace, std::vector<MappingOut>& out)
{
	out.clear();
	if (trace.empty()) return false;
	size_t pos = 0;
	size_t oldNode = g.indexToNode(trace[0].first);
	while (oldNode == 0)
	{
		pos++;
		if (pos == trace.size()) return false;
		oldNode = g.indexToNode(trace[pos].first);
	}
	if (oldNode == g.size() - 1) return false;
	out.push_back({ g.nodeIDs[oldNode], g.reverse[oldNode] ? 1 : 0, (long)(trace[pos].first - g.start(oldNode)), 0, 0 });
	Pos nodeStart = trace[pos], nodeEnd = trace[pos], beforeNode = trace[pos];
	for (; pos < trace.size(); pos++)
	{
		size_t n = g.indexToNode(trace[pos].first);
		if (n == g.size() - 1) break;
		if (n == oldNode) { nodeEnd = trace[pos]; continue; }
		out.back().from_length = (int)(nodeEnd.first - nodeStart.first + 1);
		out.back().to_length = (int)(nodeEnd.second - beforeNode.second);
		oldNode = n;
		beforeNode = nodeEnd;
		nodeStart = nodeEnd = trace[pos];
		out.push_back({ g.nodeIDs[oldNode], g.reverse[oldNode] ? 1 : 0, 0, 0, 0 });
	}
	out.back().from_length = (int)(nodeEnd.first - nodeStart.first);   // no +1 on the last mapping, GraphAligner.h:843
	out.back().to_length = (int)(nodeEnd.second - beforeNode.second);
	return true;
}

// getTraceInfoInner, GraphAligner.h:718-780
static void traceInfoInner(const Graph& g, const std::string& sequence, const std::vector<Pos>& trace, std::vector<TraceItem>& result)
{
	for (size_t i = 1; i < trace.size(); i++)
	{
		Pos np = trace[i], op = trace[i - 1];
		size_t n = g.indexToNode(np.first);
		bool diagonal = np.second != op.second;
		if (np.first == op.first)
		{
			bool selfLoop = std::find(g.out[n].begin(), g.out[n].end(), n) != g.out[n].end();
			if (!(np.second == op.second + 1 && g.len(n) == 1 && selfLoop)) diagonal = false;
		}
		TraceItem item;
		item.nodeID = g.nodeIDs[n] / 2;
		item.reverse = g.nodeIDs[n] % 2 == 1;
		item.offset = np.first - g.start(n);
		item.readpos = (size_t)np.second;
		if (np.second == op.second) item.type = 4;
		else if (np.first == op.first && !diagonal) item.type = 3;
		else item.type = characterMatch(sequence[np.second], g.seq[np.first]) ? 1 : 2;
		result.push_back(item);
	}
}

static ReadResult alignRead(const Graph& g, const std::string& sequence, const std::vector<std::tuple<int, size_t, bool>>& seeds, int b, int B)
{
	ReadResult res;
	Aligner al(g, b, B);
	std::vector<std::tuple<size_t, size_t, size_t>> tried;
	bool hasAlignment = false;
	size_t bestEstimated = 0, bestSeedPos = 0;
	std::pair<int, std::vector<Pos>> bestFw, bestBw;
	for (auto& seed : seeds)
	{
		size_t nodeIndex = g.lookup.at(std::get<0>(seed) * 2);
		size_t pos = std::get<1>(seed);
		bool already = false;
		for (auto& t : tried) already = already || (std::get<0>(t) <= pos && std::get<1>(t) >= pos && std::get<2>(t) == nodeIndex);
		if (already) continue;   // GraphAligner.h:425-429
		// getSplitAlignment, GraphAligner.h:2969-3024
		size_t forwardNode = g.lookup.at(std::get<0>(seed) * 2 + (std::get<2>(seed) ? 1 : 0));
		size_t backwardNode = g.lookup.at(std::get<0>(seed) * 2 + (std::get<2>(seed) ? 0 : 1));
		std::pair<int, std::vector<Pos>> fw { 0, {} }, bw { 0, {} };
		size_t fwSlices = 0, bwSlices = 0;
		bool bwPresent = false;
		if (pos < sequence.size() - 1)
		{
			std::string part = sequence.substr(pos);
			part.append((64 - part.size() % 64) % 64, 'N');
			Aligner::Table t = al.forward(part, forwardNode);
			fwSlices = t.correctness.size();
			if (fwSlices > 0)
			{
				fw = al.trace(t, part);
				size_t limit = sequence.size() - pos - g.overlap;   // GraphAligner.h:3063-3066
				while (!fw.second.empty() && (size_t)fw.second.back().second >= limit) fw.second.pop_back();
			}
		}
		if (pos > 0)
		{
			std::string part = reverseComplement(sequence.substr(0, pos + g.overlap));
			part.append((64 - part.size() % 64) % 64, 'N');
			Aligner::Table t = al.forward(part, backwardNode);
			bwSlices = t.correctness.size();
			if (bwSlices > 0)
			{
				bwPresent = true;
				bw = al.trace(t, part);
				while (!bw.second.empty() && (size_t)bw.second.back().second >= pos) bw.second.pop_back();   // GraphAligner.h:3086-3089
				std::reverse(bw.second.begin(), bw.second.end());                                             // reverseTrace, :3026-3037
				for (auto& p : bw.second) { p.first = g.reversePosition(p.first); p.second = (long)(pos - 1) - p.second; }
				for (auto& p : fw.second) p.second += (long)pos;   // only inside this branch, GraphAligner.h:3090-3093
			}
		}
		(void)bwPresent;
		// addAlignmentNodes, GraphAligner.h:594-634
		for (auto* tr : { &fw.second, &bw.second })
		{
			if (tr->empty()) continue;
			size_t oldNode = g.indexToNode((*tr)[0].first), startIndex = (size_t)(*tr)[0].second, endIndex = startIndex;
			for (size_t i = 1; i < tr->size(); i++)
			{
				size_t n = g.indexToNode((*tr)[i].first);
				if (n != oldNode) { tried.emplace_back(startIndex, endIndex, oldNode); startIndex = (size_t)(*tr)[i].second; oldNode = n; }
				endIndex = (size_t)(*tr)[i].second;
			}
			tried.emplace_back(startIndex, endIndex, oldNode);
		}
		size_t estimated = (fwSlices + bwSlices) * 64;
		if (!hasAlignment || estimated > bestEstimated)
		{
			bestFw = fw;
			bestBw = bw;
			bestEstimated = estimated;
			bestSeedPos = pos;
			hasAlignment = true;
		}
	}
	if (!hasAlignment) return res;
	// getTraceInfo, GraphAligner.h:690-716
	if (!bestBw.second.empty()) traceInfoInner(g, sequence, bestBw.second, res.trace);
	if (!bestBw.second.empty() && !bestFw.second.empty())
	{
		Pos p = bestFw.second[0];
		size_t n = g.indexToNode(p.first);
		res.trace.push_back({ g.nodeIDs[n] / 2, p.first - g.start(n), n % 2 == 1, (size_t)p.second, 5 });
	}
	if (!bestFw.second.empty()) traceInfoInner(g, sequence, bestFw.second, res.trace);
	std::vector<MappingOut> fwMap, bwMap;
	bool fwOk = traceToMappings(g, bestFw.second, fwMap), bwOk = traceToMappings(g, bestBw.second, bwMap);
	if (!fwOk && !bwOk) { res.trace.clear(); return res; }
	// mergeAlignments(bw, fw), GraphAligner.h:648-688
	if (!bwOk) { res.mappings = fwMap; res.score = bestFw.first; }
	else if (!fwOk) { res.mappings = bwMap; res.score = bestBw.first; }
	else
	{
		res.score = bestBw.first + bestFw.first;
		res.mappings = bwMap;
		size_t startIdx = (bwMap.back().node_id == fwMap.front().node_id && bwMap.back().is_reverse == fwMap.front().is_reverse) ? 1 : 0;
		res.mappings.insert(res.mappings.end(), fwMap.begin() + startIdx, fwMap.end());
	}
	size_t lastAligned = !bestBw.second.empty() ? (size_t)bestBw.second[0].second : bestSeedPos;
	res.qpos = (int)lastAligned;
	res.start = lastAligned;
	res.end = lastAligned + bestEstimated;
	res.failed = false;
	return res;
}

static uint64_t fnv(uint64_t h, uint64_t v)
{
	for (int i = 0; i < 8; i++) { h ^= (v >> (8 * i)) & 0xff; h *= 1099511628211ull; }
	return h;
}

int main(int argc, char** argv)
{
	if (argc < 2) { fprintf(stderr, "usage: ga_oracle <case.gacase> [--full] [--limit N]\n"); return 2; }
	bool full = false;
	size_t limit = (size_t)-1;
	for (int i = 2; i < argc; i++)
	{
		if (!strcmp(argv[i], "--full")) full = true;
		else if (!strcmp(argv[i], "--limit") && i + 1 < argc) limit = strtoull(argv[++i], nullptr, 10);
	}
	std::ifstream in(argv[1]);
	if (!in.good()) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
	Graph g;
	bool gfa = false;
	int b = 10, B = 0;
	std::vector<std::pair<long, std::string>> nodes;
	std::vector<std::tuple<long, bool, long, bool>> edges;
	struct Read { std::string name, seq; std::vector<std::tuple<int, size_t, bool>> seeds; };
	std::vector<Read> reads;
	std::string line;
	while (std::getline(in, line))
	{
		std::stringstream ss(line);
		std::string tag;
		ss >> tag;
		if (tag == "G") { std::string k; ss >> k; gfa = k == "gfa"; if (gfa) ss >> g.overlap; }
		else if (tag == "N") { long id; std::string s; ss >> id >> s; nodes.emplace_back(id, s); }
		else if (tag == "E") { long f, t; int fs, te; ss >> f >> fs >> t >> te; edges.emplace_back(f, fs != 0, t, te != 0); }
		else if (tag == "P") ss >> b >> B;
		else if (tag == "R")
		{
			Read r;
			size_t n;
			ss >> r.name >> r.seq >> n;
			for (size_t i = 0; i < n; i++)
			{
				std::getline(in, line);
				std::stringstream s2(line);
				std::string t2; int node; size_t pos; int rev;
				s2 >> t2 >> node >> pos >> rev;
				r.seeds.emplace_back(node, pos, rev != 0);
			}
			if (reads.size() < limit) reads.push_back(r);
		}
	}
	// BigraphToDigraph.cpp:27-104: forward node 2*id, reverse complement 2*id+1; an edge and its mirror
	for (auto& n : nodes)
	{
		std::string fwd = n.second, rev = reverseComplement(n.second);
		if (gfa) { fwd = fwd.substr(0, fwd.size() - g.overlap); rev = rev.substr(0, rev.size() - g.overlap); }
		g.addNode((int)(n.first * 2), fwd, false);
		g.addNode((int)(n.first * 2 + 1), rev, true);
	}
	for (auto& e : edges)
	{
		long from = std::get<0>(e), to = std::get<2>(e);
		bool fs = std::get<1>(e), te = std::get<3>(e);
		g.addEdge((int)(from * 2 + (fs ? 1 : 0)), (int)(to * 2 + (te ? 1 : 0)));
		g.addEdge((int)(to * 2 + (te ? 0 : 1)), (int)(from * 2 + (fs ? 0 : 1)));
	}
	g.finalize();
	for (auto& r : reads)
	{
		ReadResult res;
		if (!r.seeds.empty()) res = alignRead(g, r.seq, r.seeds, b, B);
		uint64_t h = 14695981039346656037ull;
		for (auto& t : res.trace)
		{
			h = fnv(h, (uint64_t)(int64_t)t.nodeID); h = fnv(h, t.offset); h = fnv(h, t.reverse ? 1 : 0); h = fnv(h, t.readpos); h = fnv(h, (uint64_t)t.type);
		}
		bool failed = res.failed;
		printf("READ %s failed=%d asserted=0 score=%d start=%zu end=%zu qpos=%d nmap=%zu ntrace=%zu th=%016llx\n", r.name.c_str(), failed ? 1 : 0, failed ? 0 : res.score,
			failed ? (size_t)0 : res.start, failed ? (size_t)0 : res.end, failed ? 0 : res.qpos, failed ? (size_t)0 : res.mappings.size(), failed ? (size_t)0 : res.trace.size(),
			(unsigned long long)(failed ? 0 : h));
		if (failed) continue;
		for (auto& m : res.mappings) printf("M %ld %d %ld %d %d\n", m.node_id, m.is_reverse, m.offset, m.from_length, m.to_length);
		if (full) for (auto& t : res.trace) printf("T %d %zu %d %zu %d\n", t.nodeID, t.offset, t.reverse ? 1 : 0, t.readpos, t.type);
	}
	return 0;
}
