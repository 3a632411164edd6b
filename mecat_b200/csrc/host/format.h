// mecat_b200/csrc/host/format.h -- text of the reference's result files, written without iostreams.
//
// Same characters as operator<<(ExtensionCandidate) (src/common/alignment.cpp:18-32) and output_m4record /
// operator<<(M4Record) (src/mecat2pw/pw_impl.cpp:509-531, alignment.cpp:58-78): tab separated decimal integers, the
// identity as a default-formatted double (what `out << double` prints: %g with 6 significant digits), '\n' line ends.
// A million M4 lines take ~0.1 s this way against ~1.4 s through std::ostream.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../../include/mecat_b200.h"

namespace mbfmt {

struct TextBuf
{
	std::string s;
	void chr(char c) { s.push_back(c); }
	void i64(int64_t v)
	{
		char tmp[24];
		int n = 0;
		uint64_t u = v < 0 ? 0 - (uint64_t)v : (uint64_t)v;
		do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
		if (v < 0) tmp[n++] = '-';
		const size_t at = s.size();
		s.resize(at + (size_t)n);
		for (int k = 0; k < n; ++k) s[at + (size_t)k] = tmp[n - 1 - k];
	}
	void dbl(double v)      // std::ostream's default floating-point format
	{
		char tmp[40];
		const int n = snprintf(tmp, sizeof tmp, "%g", v);
		s.append(tmp, (size_t)n);
	}
};

// printf("%.3f") of a float (promoted to double, as the reference's fprintf sees it, mecat2asmpw.c:944): the value is
// m x 2^e with a 24-bit m, so m x 1000 is an exact integer and the rounding (half to even on the exact value, what glibc
// does) needs no floating point.  Finite values below 2^39 only; anything else goes through snprintf.
inline void fixed3(TextBuf& b, float v)
{
	uint32_t bits;
	memcpy(&bits, &v, 4);
	const uint32_t ex = (bits >> 23) & 0xffu;
	uint64_t m = bits & 0x7fffffu;
	int e;                                     // v = m x 2^e
	if (ex == 0) e = -149; else { m |= 0x800000u; e = (int)ex - 150; }
	if (ex == 0xffu || e > 15) { char tmp[64]; const int n = snprintf(tmp, sizeof tmp, "%.3f", (double)v); b.s.append(tmp, (size_t)n); return; }
	uint64_t q = m * 1000u;                    // < 2^34
	if (e >= 0) q <<= e;
	else if (-e >= 64) q = 0;
	else {
		const int sft = -e;
		const uint64_t rem = q & ((sft == 64 ? 0 : ((uint64_t)1 << sft)) - 1), half = (uint64_t)1 << (sft - 1);
		q >>= sft;
		if (rem > half || (rem == half && (q & 1u))) ++q;
	}
	if (bits >> 31) b.chr('-');                // "-0.000" like printf
	b.i64((int64_t)(q / 1000));
	b.chr('.');
	const unsigned f = (unsigned)(q % 1000);
	b.chr((char)('0' + f / 100)); b.chr((char)('0' + f / 10 % 10)); b.chr((char)('0' + f % 10));
}

// one line of mecat2asmpw / mecat2trimpw (mecat2asmpw.c:944-945)
inline void format_asm(TextBuf& b, const mecat_asm_overlap* o, size_t n)
{
	for (size_t i = 0; i < n; ++i) {
		const mecat_asm_overlap& r = o[i];
		b.i64(r.sread); b.chr(' '); b.i64(r.qread); b.chr(' '); fixed3(b, r.score); b.s.append(" 100 0 ", 7);
		b.i64(r.sbeg); b.chr(' '); b.i64(r.send); b.chr(' '); b.i64(r.slen); b.chr(' '); b.i64(r.strand); b.chr(' ');
		b.i64(r.qbeg); b.chr(' '); b.i64(r.qend); b.chr(' '); b.i64(r.qlen); b.chr('\n');
	}
}

inline void format_candidates(TextBuf& b, const mecat_candidate* ec, size_t n)
{
	for (size_t i = 0; i < n; ++i) {
		const mecat_candidate& e = ec[i];
		b.i64(e.qid); b.chr('\t'); b.i64(e.sid); b.chr('\t'); b.i64(e.qdir); b.chr('\t'); b.i64(e.sdir); b.chr('\t');
		b.i64(e.qext); b.chr('\t'); b.i64(e.sext); b.chr('\t'); b.i64(e.score); b.chr('\t'); b.i64(e.qsize); b.chr('\t');
		b.i64(e.ssize); b.chr('\n');
	}
}

inline void format_m4(TextBuf& b, const mecat_m4* m, size_t n, bool gapped)
{
	for (size_t i = 0; i < n; ++i) {
		const mecat_m4& r = m[i];
		b.i64(r.qid); b.chr('\t'); b.i64(r.sid); b.chr('\t'); b.dbl(r.ident); b.chr('\t'); b.i64(r.vscore); b.chr('\t');
		b.i64(r.qdir); b.chr('\t'); b.i64(r.qoff); b.chr('\t'); b.i64(r.qend); b.chr('\t'); b.i64(r.qsize); b.chr('\t');
		b.i64(r.sdir); b.chr('\t'); b.i64(r.soff); b.chr('\t'); b.i64(r.send); b.chr('\t'); b.i64(r.ssize);
		if (gapped) { b.chr('\t'); b.i64(r.qext); b.chr('\t'); b.i64(r.sext); }
		b.chr('\n');
	}
}

}  // namespace mbfmt
