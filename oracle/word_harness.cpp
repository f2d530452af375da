// TEST INFRASTRUCTURE ONLY.  Word-level differential harness: the product's column primitives,
// compiled for the host from the product's own header (ga_core.cuh, -DGA_HOSTSIM), against the reference's WordSlice
// (compiled where it lies: $(REF)/WordSlice.h) on random columns:
//   ga_merge_cols   vs  WordSlice::mergeWith      (WordSlice.h:202, mergeTwoSlices :361-421, differenceMasks :512-615)
//   ga_col_value    vs  WordSlice::getValue       (WordSlice.h:223-229), every row
//   ga_vertical_merge vs mergeWith with the vertical ramp column the reference builds (GraphAligner.h:1541-1546)
// Prints one summary line; exit code 1 on the first difference (with the inputs).
//   word_harness [cases] [seed]
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <random>
#include <type_traits>
#include <utility>
#include <cuda_runtime.h>   /* oracle/hostsim stand-in: uint4 etc. */
#include "WordSlice.h"
#include "../graphaligner_b200/csrc/ga_core.cuh"

typedef WordSlice<size_t, int, uint64_t> RefSlice;

static GaCol randomColumn(std::mt19937_64& rng, int32_t base)
{
	// any disjoint (VP, VN) is a column; mix dense, sparse and run-structured words like the ones the DP produces
	GaCol c;
	uint64_t a = rng(), b = rng();
	switch (rng() % 5)
	{
		case 0: break;
		case 1: a &= rng(); b &= rng(); break;
		case 2: a |= rng(); break;
		case 3: a = ~(uint64_t)0 << (rng() % 64); b = rng() & ~a; break;
		default: a = (rng() % 3 == 0) ? 0 : a; b = (rng() % 3 == 0) ? 0 : b; break;
	}
	c.VP = a & ~b;
	c.VN = b & ~a;
	c.sbs = base + (int32_t)(rng() % 9) - 4;
	c.scoreEnd = c.sbs + (int32_t)__builtin_popcountll(c.VP) - (int32_t)__builtin_popcountll(c.VN);
	return c;
}

static RefSlice toRef(const GaCol& c) { return RefSlice(c.VP, c.VN, c.scoreEnd, c.sbs, 64, true); }

static bool same(const GaCol& mine, const RefSlice& ref) { return mine.VP == ref.VP && mine.VN == ref.VN && mine.sbs == ref.scoreBeforeStart && mine.scoreEnd == ref.scoreEnd; }

static void dump(const char* what, const GaCol& a, const GaCol& b, const GaCol& mine, const RefSlice& ref)
{
	fprintf(stderr, "%s differs\n  A  VP %016llx VN %016llx sbs %d end %d\n  B  VP %016llx VN %016llx sbs %d end %d\n  mine VP %016llx VN %016llx sbs %d end %d\n  ref  VP %016llx VN %016llx sbs %d end %d\n", what,
		(unsigned long long)a.VP, (unsigned long long)a.VN, a.sbs, a.scoreEnd, (unsigned long long)b.VP, (unsigned long long)b.VN, b.sbs, b.scoreEnd,
		(unsigned long long)mine.VP, (unsigned long long)mine.VN, mine.sbs, mine.scoreEnd, (unsigned long long)ref.VP, (unsigned long long)ref.VN, (int)ref.scoreBeforeStart, (int)ref.scoreEnd);
}

int main(int argc, char** argv)
{
	const size_t cases = argc > 1 ? strtoull(argv[1], nullptr, 10) : 200000;
	std::mt19937_64 rng(argc > 2 ? strtoull(argv[2], nullptr, 10) : 12345);
	size_t merges = 0, values = 0, verticals = 0;
	for (size_t it = 0; it < cases; it++)
	{
		const int32_t base = 100 + (int32_t)(rng() % 1000);
		const GaCol a = randomColumn(rng, base), b = randomColumn(rng, base);
		// element-wise minimum of two columns
		const GaCol m = ga_merge_cols(a, b);
		const RefSlice r = toRef(a).mergeWith(toRef(b));
		if (!same(m, r)) { dump("merge", a, b, m, r); return 1; }
		merges++;
		// cell values
		const RefSlice ra = toRef(a);
		for (int row = 0; row < 64; row++)
		{
			if (ga_col_value(a.VP, a.VN, a.sbs, row) != ra.getValue(row)) { fprintf(stderr, "value differs at row %d\n", row); return 1; }
			values++;
		}
		// minimum with the vertical ramp from the previous slice's end score (GraphAligner.h:1541-1546): the reference merges
		// with WordSlice(all VP, no VN, top + 64, top); ga_vertical_merge is called where top < sbs
		const int32_t top = a.sbs - 1 - (int32_t)(rng() % 4);
		GaCol v = a;
		ga_vertical_merge(v, top);
		const RefSlice rv = toRef(a).mergeWith(RefSlice(~(uint64_t)0, 0, top + 64, top, 64, true));
		if (!same(v, rv)) { GaCol ramp; ramp.VP = ~(uint64_t)0; ramp.VN = 0; ramp.sbs = top; ramp.scoreEnd = top + 64; dump("vertical merge", a, ramp, v, rv); return 1; }
		verticals++;
	}
	printf("word harness: %zu merges, %zu cell values, %zu vertical merges identical\n", merges, values, verticals);
	return 0;
}
