// mecat_b200/csrc/host/format.h -- text of the reference's result files, written without iostreams.
//
// Same characters as operator<<(ExtensionCandidate) (src/common/alignment.cpp:18-32) and output_m4record /
// operator<<(M4Record) (src/mecat2pw/pw_impl.cpp:509-531, alignment.cpp:58-78): tab separated decimal integers, the
// identity as a default-formatted double (what `out << double` prints: %g with 6 significant digits), '\n' line ends.
// A million M4 lines take ~0.1 s this way against ~1.4 s through std::ostream.
#pragma once
#include <stdint.h>
#include <stdio.h>

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
