// mecat_b200/csrc/m4_core.cuh -- per-read bodies of the record assembly and of the result text.
//
// After the extensions the reference turns a read's alignments into M4 records, orders them with std::sort and drops
// the contained ones, and prints every record as a line of tab-separated decimal numbers:
//   fill_m4record, append_m4v, check_records_containment, CmpM4RecordByQidAndOvlpSize   src/mecat2pw/pw_impl.cpp:467-610
//   operator<<(ExtensionCandidate), operator<<(M4Record)                                 src/common/alignment.cpp:18-32,58-78
// Two things make "the same bytes" more than arithmetic:
//   * std::sort is not stable, and two candidates of one read pair often extend to the very same alignment (equal keys)
//     while differing in score and extension point, which are printed; which of them survives the containment filter
//     is decided by where std::sort happened to leave them.  std_sort() below is that algorithm -- introsort as
//     libstdc++ implements it (median-of-three to first, unguarded partition, depth limit 2 floor(log2 n) with heap
//     sort behind it, final insertion sort with threshold 16) -- so the permutation is the library's, comparison by
//     comparison (checked against the real std::sort in the CPU suite, including adversarial inputs that reach the
//     heap sort).  The parallel mode the reference is built with (-D_GLIBCXX_PARALLEL) sorts fewer than 1 000 elements
//     with this sequential algorithm.
//   * `out << double` prints the identity with six significant digits, correctly rounded from the exact binary value
//     (printf's %g); fmt_g6() does that in 128-bit integer arithmetic.
// Integer / byte code shared by the CUDA backend (records.cu) and the host harness of the CPU test-suite.
#pragma once
#include <stdint.h>

#include "../../include/mecat_b200.h"

#if defined(__CUDACC__)
#define M4_HD __host__ __device__ __forceinline__
#define M4_HDN __host__ __device__
#else
#define M4_HD inline
#define M4_HDN inline
#endif

namespace mbm4 {

// ------------------------------------------------------------------------------------------ std::sort
struct SortItem { uint64_t key; int32_t idx; int32_t pad; };       // a < b  <=>  a.key < b.key

M4_HD void swap_items(SortItem* a, int i, int j) { const SortItem t = a[i]; a[i] = a[j]; a[j] = t; }

M4_HD void unguarded_linear_insert(SortItem* a, int last)
{
	const SortItem val = a[last];
	int next = last - 1;
	while (val.key < a[next].key) { a[last] = a[next]; last = next; --next; }
	a[last] = val;
}
M4_HD void insertion_sort(SortItem* a, int first, int last)
{
	if (first == last) return;
	for (int i = first + 1; i != last; ++i) {
		if (a[i].key < a[first].key) {
			const SortItem val = a[i];
			for (int j = i; j > first; --j) a[j] = a[j - 1];
			a[first] = val;
		} else unguarded_linear_insert(a, i);
	}
}
M4_HD void push_heap(SortItem* a, int first, int hole, int top, const SortItem& value)
{
	int parent = (hole - 1) / 2;
	while (hole > top && a[first + parent].key < value.key) {
		a[first + hole] = a[first + parent];
		hole = parent;
		parent = (hole - 1) / 2;
	}
	a[first + hole] = value;
}
M4_HD void adjust_heap(SortItem* a, int first, int hole, int len, const SortItem& value)
{
	const int top = hole;
	int child = hole;
	while (child < (len - 1) / 2) {
		child = 2 * (child + 1);
		if (a[first + child].key < a[first + child - 1].key) --child;
		a[first + hole] = a[first + child];
		hole = child;
	}
	if ((len & 1) == 0 && child == (len - 2) / 2) {
		child = 2 * (child + 1);
		a[first + hole] = a[first + child - 1];
		hole = child - 1;
	}
	push_heap(a, first, hole, top, value);
}
M4_HD void heap_sort(SortItem* a, int first, int last)      // std::partial_sort(first, last, last): make_heap + sort_heap
{
	const int len = last - first;
	if (len >= 2) {
		for (int parent = (len - 2) / 2;; --parent) {
			const SortItem value = a[first + parent];
			adjust_heap(a, first, parent, len, value);
			if (parent == 0) break;
		}
	}
	for (int end = last; end - first > 1;) {
		--end;
		const SortItem value = a[end];
		a[end] = a[first];
		adjust_heap(a, first, 0, end - first, value);
	}
}
// one partitioning step of the introsort loop: median of (first+1, mid, last-1) to first, unguarded partition around it
M4_HD int partition_pivot(SortItem* a, int first, int last)
{
	const int mid = first + (last - first) / 2;
	const int x = first + 1, y = mid, z = last - 1;
	if (a[x].key < a[y].key) {
		if (a[y].key < a[z].key) swap_items(a, first, y);
		else if (a[x].key < a[z].key) swap_items(a, first, z);
		else swap_items(a, first, x);
	} else if (a[x].key < a[z].key) swap_items(a, first, x);
	else if (a[y].key < a[z].key) swap_items(a, first, z);
	else swap_items(a, first, y);
	int lo = first + 1, hi = last;
	for (;;) {
		while (a[lo].key < a[first].key) ++lo;
		--hi;
		while (a[first].key < a[hi].key) --hi;
		if (!(lo < hi)) return lo;
		swap_items(a, lo, hi);
		++lo;
	}
}
// std::sort(a, a + n).  The two halves a partition leaves are disjoint, so the order they are finished in does not
// matter: the recursion of the library on the right half becomes an explicit stack.
M4_HDN void std_sort(SortItem* a, int n, int* heap_sorts = nullptr /* test hook: ranges that fell through to the heap sort */)
{
	if (n <= 1) return;
	int lg = 0;
	while ((n >> (lg + 1)) > 0) ++lg;
	int sf[64], sl[64], sd[64], sp = 0;
	sf[0] = 0; sl[0] = n; sd[0] = 2 * lg; sp = 1;
	while (sp > 0) {
		--sp;
		const int first = sf[sp];
		int last = sl[sp], depth = sd[sp];
		while (last - first > 16) {
			if (depth == 0) { heap_sort(a, first, last); if (heap_sorts) ++*heap_sorts; break; }
			--depth;
			const int cut = partition_pivot(a, first, last);
			sf[sp] = cut; sl[sp] = last; sd[sp] = depth; ++sp;      // the library recurses into [cut, last) here
			last = cut;
		}
	}
	if (n > 16) {
		insertion_sort(a, 0, 16);
		for (int i = 16; i != n; ++i) unguarded_linear_insert(a, i);
	} else insertion_sort(a, 0, n);
}

// ------------------------------------------------------------------------------------------ A12
// fill_m4record, pw_impl.cpp:467-506: the index-side read becomes qid, reverse-strand coordinates are flipped.
M4_HD mecat_m4 make_m4(int64_t ref_read_id, int64_t ref_read_size, int64_t query_id, int64_t query_size, int qstrand, int task_qstart,
                       int task_sstart, const mecat_extend_result& R, int score)
{
	mecat_m4 m;
	m.qid = ref_read_id; m.sid = query_id; m.ident = R.ident; m.vscore = score; m.qdir = 0;
	m.qoff = R.sstart; m.qend = R.send; m.qsize = ref_read_size;
	m.pad_ = 0;
	m.ssize = query_size; m.qext = task_sstart;
	if (!qstrand) { m.sdir = 0; m.soff = R.qstart; m.send = R.qend; m.sext = task_qstart; }
	else { m.sdir = 1; m.soff = query_size - R.qend; m.send = query_size - R.qstart; m.sext = query_size - 1 - task_qstart; }
	return m;
}
// CmpM4RecordByQidAndOvlpSize, pw_impl.cpp:539-548, as one unsigned key: qid ascending, then the shorter of the two
// aligned spans descending
M4_HD uint64_t m4_key(int64_t qid, int64_t qspan, int64_t sspan)
{
	const int64_t o = qspan < sspan ? qspan : sspan;
	return ((uint64_t)qid << 32) | (uint32_t)(0x7FFFFFFF - (int32_t)o);
}
// check_records_containment, pw_impl.cpp:550-574: is b inside a (same pair, same strand, 100 bases of slack)
M4_HD bool m4_contained(const mecat_m4& a, const mecat_m4& b)
{
	return a.sdir == b.sdir && b.qoff + 100 >= a.qoff && b.qend - 100 <= a.qend && b.soff + 100 >= a.soff && b.send - 100 <= a.send;
}

// ------------------------------------------------------------------------------------------ text
M4_HD int fmt_i64(char* out, int64_t v)
{
	char tmp[24];
	int n = 0;
	uint64_t u = v < 0 ? 0 - (uint64_t)v : (uint64_t)v;
	do { tmp[n++] = (char)('0' + (int)(u % 10)); u /= 10; } while (u);
	if (v < 0) tmp[n++] = '-';
	for (int k = 0; k < n; ++k) out[k] = tmp[n - 1 - k];
	return n;
}

// printf("%g", v) for 0 <= v < 1e6 (the identity column is 100 * matches / columns); returns the length, -1 outside
// that range.  The six digits are round-half-even of the exact value v * 10^k, v = f * 2^e.
M4_HDN int fmt_g6(char* out, double v)
{
	if (v == 0.0) { out[0] = '0'; return 1; }
	if (!(v >= 1e-9 && v < 1e6)) return -1;
	union { double d; uint64_t u; } cv;
	cv.d = v;
	const int e2 = (int)((cv.u >> 52) & 0x7ff);
	uint64_t f = cv.u & ((1ull << 52) - 1);
	int e;
	if (e2 == 0) e = -1074; else { f |= 1ull << 52; e = e2 - 1075; }
	// decimal exponent: a first guess from the binary one, corrected by the digit count below
	int X = (int)(((e + 52) * 1233) >> 12);            // floor(log10(2^(e+52))) up to one
	unsigned long long digits = 0;
	for (int tries = 0; tries < 4; ++tries) {
		const int k = 5 - X;                           // 0 <= k <= 15 in range
		unsigned __int128 num = (unsigned __int128)f;
		for (int i = 0; i < k; ++i) num *= 10u;
		for (int i = k; i < 0; ++i) num /= 10u;        // not reached in range (k >= 0)
		const int shift = -e;                          // e < 0 in range
		unsigned __int128 q = num >> shift;
		const unsigned __int128 rem = num - (q << shift), half = (unsigned __int128)1 << (shift - 1);
		if (rem > half || (rem == half && (q & 1))) ++q;
		digits = (unsigned long long)q;
		if (digits < 100000ull) { --X; continue; }
		if (digits >= 1000000ull) { ++X; continue; }
		break;
	}
	char d[6];
	for (int i = 5; i >= 0; --i) { d[i] = (char)('0' + (int)(digits % 10)); digits /= 10; }
	int nd = 6;
	while (nd > 1 && d[nd - 1] == '0') --nd;           // %g drops trailing zeros
	int n = 0;
	if (X < -4 || X >= 6) {                            // exponent style d.ddddde-XX
		out[n++] = d[0];
		if (nd > 1) { out[n++] = '.'; for (int i = 1; i < nd; ++i) out[n++] = d[i]; }
		out[n++] = 'e';
		int ax = X;
		if (ax < 0) { out[n++] = '-'; ax = -ax; } else out[n++] = '+';
		if (ax >= 100) { out[n++] = (char)('0' + ax / 100); ax %= 100; }
		out[n++] = (char)('0' + ax / 10); out[n++] = (char)('0' + ax % 10);
		return n;
	}
	if (X >= 0) {
		for (int i = 0; i <= X; ++i) out[n++] = i < 6 ? d[i] : '0';
		if (nd > X + 1) { out[n++] = '.'; for (int i = X + 1; i < nd; ++i) out[n++] = d[i]; }
		return n;
	}
	out[n++] = '0'; out[n++] = '.';
	for (int i = 0; i < -X - 1; ++i) out[n++] = '0';
	for (int i = 0; i < nd; ++i) out[n++] = d[i];
	return n;
}

constexpr int LINE_CAP = 320;      // 14 fields of at most 20 characters, separators

// operator<<(ExtensionCandidate), alignment.cpp:18-32
M4_HD int line_candidate(char* b, const mecat_candidate& e)
{
	int n = 0;
	n += fmt_i64(b + n, e.qid); b[n++] = '\t'; n += fmt_i64(b + n, e.sid); b[n++] = '\t'; n += fmt_i64(b + n, e.qdir); b[n++] = '\t';
	n += fmt_i64(b + n, e.sdir); b[n++] = '\t'; n += fmt_i64(b + n, e.qext); b[n++] = '\t'; n += fmt_i64(b + n, e.sext); b[n++] = '\t';
	n += fmt_i64(b + n, e.score); b[n++] = '\t'; n += fmt_i64(b + n, e.qsize); b[n++] = '\t'; n += fmt_i64(b + n, e.ssize); b[n++] = '\n';
	return n;
}
// output_m4record / operator<<(M4Record), pw_impl.cpp:509-531, alignment.cpp:58-78; -1 when the identity is out of range
M4_HD int line_m4(char* b, const mecat_m4& r, bool gapped)
{
	int n = 0;
	n += fmt_i64(b + n, r.qid); b[n++] = '\t'; n += fmt_i64(b + n, r.sid); b[n++] = '\t';
	const int g = fmt_g6(b + n, r.ident);
	if (g < 0) return -1;
	n += g; b[n++] = '\t';
	n += fmt_i64(b + n, r.vscore); b[n++] = '\t'; n += fmt_i64(b + n, r.qdir); b[n++] = '\t'; n += fmt_i64(b + n, r.qoff); b[n++] = '\t';
	n += fmt_i64(b + n, r.qend); b[n++] = '\t'; n += fmt_i64(b + n, r.qsize); b[n++] = '\t'; n += fmt_i64(b + n, r.sdir); b[n++] = '\t';
	n += fmt_i64(b + n, r.soff); b[n++] = '\t'; n += fmt_i64(b + n, r.send); b[n++] = '\t'; n += fmt_i64(b + n, r.ssize);
	if (gapped) { b[n++] = '\t'; n += fmt_i64(b + n, r.qext); b[n++] = '\t'; n += fmt_i64(b + n, r.sext); }
	b[n++] = '\n';
	return n;
}

}  // namespace mbm4
