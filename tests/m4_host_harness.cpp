// tests/m4_host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The product's record-assembly and text bodies (mecat_b200/csrc/m4_core.cuh) on the host, next to the library
// functions they must reproduce: std_sort against the real std::sort (same permutation, ties included, also on
// adversarial inputs that drive the library into its heap sort), fmt_g6 against printf("%g"), the integer and line
// formatters against the host formatter of the drivers.  Compiled by tests/util.py; never part of the product library.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "../mecat_b200/csrc/m4_core.cuh"
#include "../mecat_b200/csrc/host/format.h"

namespace {

struct Rec { uint64_t key; int idx; char payload[92]; };      // as fat as an M4 record: std::sort moves whole records
struct RecLess { bool operator()(const Rec& a, const Rec& b) const { return a.key < b.key; } };

// 0 = same permutation
int compare_once(const std::vector<uint64_t>& keys)
{
	const int n = (int)keys.size();
	std::vector<Rec> lib((size_t)n);
	std::vector<mbm4::SortItem> mine((size_t)n);
	for (int i = 0; i < n; ++i) { lib[(size_t)i].key = keys[(size_t)i]; lib[(size_t)i].idx = i; mine[(size_t)i].key = keys[(size_t)i]; mine[(size_t)i].idx = i; mine[(size_t)i].pad = 0; }
	std::sort(lib.begin(), lib.end(), RecLess());
	mbm4::std_sort(mine.data(), n);
	for (int i = 0; i < n; ++i) if (lib[(size_t)i].idx != mine[(size_t)i].idx) return 1;
	return 0;
}

// McIlroy's adversary ("A killer adversary for quicksort"): keys are decided while the library's own std::sort runs, so
// that its partitions are as lopsided as they can be; the frozen keys then drive std::sort to its depth limit.
struct Adversary
{
	std::vector<int> val; int nsolid = 0, candidate = 0, gas;
	explicit Adversary(int n) : val((size_t)n), gas(n - 1) { for (int& v : val) v = gas; }
	bool less(int x, int y)
	{
		if (val[(size_t)x] == gas && val[(size_t)y] == gas) { if (x == candidate) val[(size_t)x] = nsolid++; else val[(size_t)y] = nsolid++; }
		if (val[(size_t)x] == gas) candidate = x; else if (val[(size_t)y] == gas) candidate = y;
		return val[(size_t)x] < val[(size_t)y];
	}
};

}  // namespace

extern "C" {

// random key arrays of every size up to max_n with few distinct values (ties everywhere); returns the number of
// arrays whose permutation differs from the library's
long mh_sort_random(int max_n, int rounds, unsigned seed)
{
	std::mt19937_64 rng(seed);
	long bad = 0;
	for (int r = 0; r < rounds; ++r)
		for (int n = 0; n <= max_n; ++n) {
			const int distinct = 1 + (int)(rng() % (uint64_t)(n < 2 ? 2 : (r % 3 == 0 ? 3 : r % 3 == 1 ? n / 2 + 1 : 4 * n)));
			std::vector<uint64_t> keys((size_t)n);
			for (auto& k : keys) k = rng() % (uint64_t)distinct;
			if (r % 5 == 0) std::sort(keys.begin(), keys.end());
			if (r % 7 == 0) std::sort(keys.rbegin(), keys.rend());
			bad += compare_once(keys);
		}
	return bad;
}

// adversarial inputs of n keys; *heap_used tells whether the product's sort reached its heap sort on them
long mh_sort_adversary(int n, int* heap_used)
{
	Adversary adv(n);
	std::vector<int> ptr((size_t)n);
	for (int i = 0; i < n; ++i) ptr[(size_t)i] = i;
	std::sort(ptr.begin(), ptr.end(), [&](int x, int y) { return adv.less(x, y); });
	std::vector<uint64_t> keys((size_t)n);
	for (int i = 0; i < n; ++i) keys[(size_t)i] = (uint64_t)adv.val[(size_t)i];
	if (heap_used) {
		std::vector<mbm4::SortItem> a((size_t)n);
		for (int i = 0; i < n; ++i) { a[(size_t)i].key = keys[(size_t)i]; a[(size_t)i].idx = i; a[(size_t)i].pad = 0; }
		*heap_used = 0;
		mbm4::std_sort(a.data(), n, heap_used);
	}
	return compare_once(keys);
}

// fmt_g6 against printf("%g"): returns the number of values that differ (first one reported in msg)
long mh_fmt_ratios(int max_n, int step, char* msg, int cap)
{
	long bad = 0;
	char a[64], b[64];
	for (int n = 1; n <= max_n; n += step)
		for (int m = 0; m <= n; m += (n > 4000 ? 1 + n / 997 : 1)) {
			const double v = 100.0 * m / n;
			const int la = mbm4::fmt_g6(a, v);
			const int lb = snprintf(b, sizeof b, "%g", v);
			if (la != lb || memcmp(a, b, (size_t)lb)) { if (!bad++) snprintf(msg, (size_t)cap, "%d/%d: %.*s vs %s", m, n, la < 0 ? 0 : la, a, b); }
		}
	return bad;
}
long mh_fmt_values(const double* v, long n, char* msg, int cap)
{
	long bad = 0;
	char a[64], b[64];
	for (long i = 0; i < n; ++i) {
		const int la = mbm4::fmt_g6(a, v[i]);
		const int lb = snprintf(b, sizeof b, "%g", v[i]);
		if (la != lb || memcmp(a, b, (size_t)lb)) { if (!bad++) snprintf(msg, (size_t)cap, "%.17g: %.*s vs %s", v[i], la < 0 ? 0 : la, a, b); }
	}
	return bad;
}

// whole lines against the drivers' host formatter
long mh_lines(const mecat_candidate* ec, long nec, const mecat_m4* m4, long nm4)
{
	long bad = 0;
	char line[mbm4::LINE_CAP];
	for (long i = 0; i < nec; ++i) {
		mbfmt::TextBuf b;
		mbfmt::format_candidates(b, ec + i, 1);
		const int l = mbm4::line_candidate(line, ec[i]);
		bad += !(l == (int)b.s.size() && !memcmp(line, b.s.data(), (size_t)l));
	}
	for (long i = 0; i < nm4; ++i)
		for (int g = 0; g < 2; ++g) {
			mbfmt::TextBuf b;
			mbfmt::format_m4(b, m4 + i, 1, g != 0);
			const int l = mbm4::line_m4(line, m4[i], g != 0);
			bad += !(l == (int)b.s.size() && !memcmp(line, b.s.data(), (size_t)l));
		}
	return bad;
}

}  // extern "C"
