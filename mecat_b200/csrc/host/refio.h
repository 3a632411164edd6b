// mecat_b200/csrc/host/refio.h -- host side of mecat2ref: input parsing, 2-bit packing, result text.
//
// Follows the reference's own readers, quirks included, so that the same files give the same records:
//   reference genome   creat_ref_index, src/mecat2ref/mecat2ref_impl_large.cpp:153-196 (name = header up to the first blank,
//                      letters above 'Z' upper-cased, every character except line ends is a base, sequences concatenated)
//   reads              chang_fastqfile, src/mecat2ref/mecat2ref.cpp:192-248 (FASTA reads are numbered from 0, FASTQ reads
//                      from 1; bases are kept as written)
//   strands            reference_mapping, mecat2ref_impl_large.cpp:355-400 (the reverse strand complements upper-case ACGT
//                      only), transnum_buchang :64-90 (only upper-case ACGT seeds), extract_sequences
//                      mecat2ref_aux.cpp:86-121 (either case aligns, any other letter aligns as A)
//   result text        print_ref_result / print_m4_result, src/mecat2ref/output.cpp:8-88; output_query_results and
//                      get_chr_id, mecat2ref.cpp:280-356
// Used by the mecat2ref driver and by the host harness of the CPU test-suite.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../../include/mecat_b200.h"

namespace refio {

inline int code_ci(unsigned char c)       // A0 C1 G2 T3 in either case, -1 for anything else
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': return 3;
	}
	return -1;
}
inline bool upper_acgt(unsigned char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

struct Packer          // the reference's volume layout: base i in byte i >> 2 at shift ((~i) & 3) << 1
{
	std::vector<uint8_t> pac;
	int64_t n = 0;
	void push(int code)
	{
		if ((n & 3) == 0) pac.push_back(0);
		pac.back() |= (uint8_t)(code << (((~n) & 3) << 1));
		++n;
	}
};

struct Chr { int64_t start = 0, size = 0; std::string name; };

struct Genome
{
	std::vector<Chr> chr;
	Packer seq;                          // all sequences concatenated, letters other than ACGT packed as A
	std::vector<int64_t> runs;           // {start, length} of every maximal run of ACGT: no k-mer of the index spans another letter
	mecat_ref_genome view() const
	{
		mecat_ref_genome g;
		g.num_bases = seq.n; g.pac = seq.pac.data(); g.num_runs = (int32_t)(runs.size() / 2); g.run_start_len = runs.data();
		return g;
	}
};

inline bool read_file(const char* path, std::string& all)
{
	FILE* f = fopen(path, "rb");
	if (!f) return false;
	char buf[1 << 16];
	size_t r;
	while ((r = fread(buf, 1, sizeof buf, f)) > 0) all.append(buf, r);
	fclose(f);
	return true;
}

inline bool load_genome(const char* path, Genome& G, std::string& err)
{
	std::string all;
	if (!read_file(path, all)) { err = std::string("cannot open ") + path; return false; }
	int64_t run_start = -1;
	size_t i = 0;
	while (i < all.size()) {
		const unsigned char ch = (unsigned char)all[i];
		if (ch == '>') {
			size_t e = all.find('\n', i);
			if (e == std::string::npos) e = all.size();
			size_t k = i + 1;
			while (k < e && all[k] != ' ' && all[k] != '\t') ++k;
			if (!G.chr.empty()) G.chr.back().size = G.seq.n - G.chr.back().start;
			Chr c; c.start = G.seq.n; c.name = all.substr(i + 1, k - i - 1);
			G.chr.push_back(c);
			i = e;
			continue;
		}
		++i;
		if (ch == '\n' || ch == '\r') continue;
		const unsigned char up = ch > 'Z' ? (unsigned char)toupper(ch) : ch;
		const bool good = upper_acgt(up);
		if (good && run_start < 0) run_start = G.seq.n;
		if (!good && run_start >= 0) { G.runs.push_back(run_start); G.runs.push_back(G.seq.n - run_start); run_start = -1; }
		G.seq.push(good ? code_ci(up) : 0);
	}
	if (run_start >= 0) { G.runs.push_back(run_start); G.runs.push_back(G.seq.n - run_start); }
	if (!G.chr.empty()) G.chr.back().size = G.seq.n - G.chr.back().start;
	if (G.chr.empty()) { err = std::string("no sequence in ") + path; return false; }
	return true;
}

struct Reads
{
	std::vector<int32_t> name;           // the number the reference prints for the read
	std::vector<std::string> seq;
};

inline bool load_reads(const char* path, Reads& R, std::string& err)
{
	std::string all;
	if (!read_file(path, all)) { err = std::string("cannot open ") + path; return false; }
	if (all.empty()) return true;
	if (all[0] == '>') {
		size_t i = 0;
		int next = 0;
		while (i < all.size()) {
			if (all[i] == '>') {
				while (i < all.size() && all[i] != '\n') ++i;
				R.name.push_back(next++);
				R.seq.push_back(std::string());
			} else {
				if (all[i] != '\n' && all[i] != '\r') R.seq.back().push_back(all[i]);
				++i;
			}
		}
	} else {
		std::vector<std::string> lines;
		size_t i = 0;
		while (i < all.size()) {
			const size_t e = all.find('\n', i);
			std::string l = all.substr(i, e == std::string::npos ? std::string::npos : e - i);
			if (!l.empty() && l[l.size() - 1] == '\r') l.erase(l.size() - 1);
			lines.push_back(l);
			if (e == std::string::npos) break;
			i = e + 1;
		}
		int next = 0;
		for (size_t k = 0; k + 3 < lines.size(); k += 4) { R.name.push_back(++next); R.seq.push_back(lines[k + 1]); }
	}
	return true;
}

// Reads [first, first + count) packed as one volume.  A read made of upper-case ACGT only is packed once (its reverse
// strand is the reverse complement of the packed bases); any other read also gets its reverse strand packed explicitly,
// built the way the reference builds it.
struct ReadBatch
{
	Packer bases;
	std::vector<int32_t> offsz, len, fread, rread, rrc;
	std::vector<int64_t> bad;
	mecat_volume vol;
	mecat_ref_reads view()
	{
		vol.num_reads = (int32_t)(offsz.size() / 2); vol.num_bases = (int32_t)bases.n; vol.start_read_id = 0;
		vol.offset_size = offsz.data(); vol.pac = bases.pac.data();
		mecat_ref_reads r;
		r.num_reads = (int32_t)len.size(); r.vol = &vol; r.read_len = len.data(); r.fwd_read = fread.data(); r.rev_read = rread.data();
		r.rev_is_rc = rrc.data(); r.num_bad = (int64_t)bad.size(); r.bad = bad.data();
		return r;
	}
	int32_t add_sequence(const std::string& s)
	{
		const int32_t id = (int32_t)(offsz.size() / 2);
		offsz.push_back((int32_t)bases.n); offsz.push_back((int32_t)s.size());
		for (size_t i = 0; i < s.size(); ++i) {
			const int c = code_ci((unsigned char)s[i]);
			if (!upper_acgt((unsigned char)s[i])) bad.push_back(bases.n);
			bases.push(c < 0 ? 0 : c);
		}
		bases.push(0);      // one pad base between reads, like the reference's volumes
		return id;
	}
	void add_read(const std::string& s)
	{
		bool plain = true;
		for (size_t i = 0; i < s.size() && plain; ++i) plain = upper_acgt((unsigned char)s[i]);
		len.push_back((int32_t)s.size());
		const int32_t f = add_sequence(s);
		fread.push_back(f);
		if (plain) { rread.push_back(f); rrc.push_back(1); return; }
		std::string r(s.rbegin(), s.rend());
		for (char& c : r) c = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
		rread.push_back(add_sequence(r)); rrc.push_back(0);
	}
	int64_t packed_bases() const { return bases.n; }
};

inline int chr_of(const std::vector<Chr>& chr, int64_t offset)      // get_chr_id, mecat2ref.cpp:280-298
{
	const int n = (int)chr.size();
	int left = 0, right = n, mid = 0;
	while (left < right) {
		mid = (left + right) >> 1;
		if (offset >= chr[(size_t)mid].start) {
			if (mid == n - 1) break;
			if (offset < chr[(size_t)mid + 1].start) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}

// format 0 = ref (header line + the two alignment strings), 1 = m4.  `names`: the printed number of read r.
inline void format_results(std::string& out, const Genome& G, const std::vector<int32_t>& names, int32_t first_read, const mecat_ref_result* recs, size_t n,
                           const char* qstr, const char* sstr, int format)
{
	char line[1400];
	for (size_t i = 0; i < n; ++i) {
		const mecat_ref_result& r = recs[i];
		const Chr& c = G.chr[(size_t)chr_of(G.chr, r.sb)];
		int qb = r.qb, qe = r.qe;
		if (r.dir) { qb = r.qs - r.qe; qe = r.qs - r.qb; }
		const int id = names[(size_t)(first_read + r.read)];
		if (format == 0) {
			snprintf(line, sizeof line, "%d\t%s\t%c\t%d\t%d\t%d\t%d\t%ld\t%ld\t%ld\n", id, c.name.c_str(), r.dir ? 'R' : 'F', r.vscore, qb, qe, r.qs,
			         (long)(r.sb - c.start), (long)(r.se - c.start), (long)c.size);
			out += line;
			out.append(qstr + r.str_offset, (size_t)r.columns); out += '\n';
			out.append(sstr + r.str_offset, (size_t)r.columns); out += '\n';
		} else {
			double ident = (double)r.matches;        // print_m4_result counts the equal columns of the two strings
			ident = ident / (double)r.columns;
			ident *= 100.0;
			snprintf(line, sizeof line, "%d\t%s\t%.4f\t%d\t%d\t%d\t%d\t%d\t0\t%ld\t%ld\t%ld\n", id, c.name.c_str(), ident, r.vscore, r.dir ? 1 : 0, qb, qe,
			         r.qs, (long)(r.sb - c.start), (long)(r.se - c.start), (long)c.size);
			out += line;
		}
	}
}

}  // namespace refio
