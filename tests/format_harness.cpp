// tests/format_harness.cpp -- TEST INFRASTRUCTURE ONLY: the drivers' text formatting (mecat_b200/csrc/host/format.h)
// next to the std::ostream formatting the reference uses, so the CPU suite can compare them character by character.
#include <string.h>

#include <sstream>

#include "../mecat_b200/csrc/host/format.h"

static size_t give(const std::string& s, char* out, size_t cap)
{
	if (s.size() <= cap) memcpy(out, s.data(), s.size());
	return s.size();
}

extern "C" {

size_t harness_format_m4(const mecat_m4* m, size_t n, int gapped, char* out, size_t cap)
{
	mbfmt::TextBuf b;
	mbfmt::format_m4(b, m, n, gapped != 0);
	return give(b.s, out, cap);
}

size_t harness_ostream_m4(const mecat_m4* m, size_t n, int gapped, char* out, size_t cap)   // operator<<(M4Record), alignment.cpp:58-78
{
	std::ostringstream os;
	for (size_t i = 0; i < n; ++i) {
		const mecat_m4& r = m[i];
		os << r.qid << '\t' << r.sid << '\t' << r.ident << '\t' << r.vscore << '\t' << r.qdir << '\t' << r.qoff << '\t' << r.qend << '\t'
		   << r.qsize << '\t' << r.sdir << '\t' << r.soff << '\t' << r.send << '\t' << r.ssize;
		if (gapped) os << '\t' << r.qext << '\t' << r.sext;
		os << "\n";
	}
	return give(os.str(), out, cap);
}

size_t harness_format_can(const mecat_candidate* e, size_t n, char* out, size_t cap)
{
	mbfmt::TextBuf b;
	mbfmt::format_candidates(b, e, n);
	return give(b.s, out, cap);
}

size_t harness_ostream_can(const mecat_candidate* ec, size_t n, char* out, size_t cap)      // operator<<(ExtensionCandidate), alignment.cpp:18-32
{
	std::ostringstream os;
	for (size_t i = 0; i < n; ++i) {
		const mecat_candidate& e = ec[i];
		os << e.qid << '\t' << e.sid << '\t' << e.qdir << '\t' << e.sdir << '\t' << e.qext << '\t' << e.sext << '\t' << e.score << '\t'
		   << e.qsize << '\t' << e.ssize << std::endl;
	}
	return give(os.str(), out, cap);
}

}  // extern "C"

// printf("%.3f") of a float by integer arithmetic (format.h fixed3): returns the length written
extern "C" int fh_fixed3(float v, char* out, int cap)
{
	mbfmt::TextBuf b;
	mbfmt::fixed3(b, v);
	const int n = (int)b.s.size() < cap - 1 ? (int)b.s.size() : cap - 1;
	memcpy(out, b.s.data(), (size_t)n);
	out[n] = 0;
	return n;
}
