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
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <deque>
#include <string>
#include <thread>
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

struct Packed          // 2 bits per base in the reference's volume layout: base i in byte i >> 2 at shift ((~i) & 3) << 1
{
	std::vector<uint8_t> pac;
	int64_t n = 0;
};

struct Chr { int64_t start = 0, size = 0; std::string name; };

struct Genome
{
	std::vector<Chr> chr;
	Packed seq;                          // all sequences concatenated, letters other than ACGT packed as A
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

struct MappedFile      // the input file, mapped read-only (falls back to reading it for non-regular files)
{
	const char* p = NULL;
	size_t n = 0;
	std::string copy;
	bool mapped = false;
	bool open(const char* path)
	{
		const int fd = ::open(path, O_RDONLY);
		if (fd < 0) return false;
		struct stat st;
		if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
			void* m = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
			if (m != MAP_FAILED) { p = (const char*)m; n = (size_t)st.st_size; mapped = true; madvise(m, n, MADV_SEQUENTIAL); }
		}
		::close(fd);
		if (mapped) return true;
		if (!read_file(path, copy)) return false;
		p = copy.data(); n = copy.size();
		return true;
	}
	~MappedFile() { if (mapped) munmap((void*)p, n); }
};

inline bool load_genome(const char* path, Genome& G, std::string& err)
{
	MappedFile F;
	if (!F.open(path)) { err = std::string("cannot open ") + path; return false; }
	// per character: bits 0-1 = code after upper-casing, bit 7 = not ACGT (packs as A, ends a run), bit 6 = line end (skipped)
	uint8_t T[256];
	for (int c = 0; c < 256; ++c) {
		const unsigned char up = c > 'Z' ? (unsigned char)toupper(c) : (unsigned char)c;
		T[c] = upper_acgt(up) ? (uint8_t)code_ci(up) : (uint8_t)0x80;
	}
	T[(unsigned char)'\n'] = 0x40; T[(unsigned char)'\r'] = 0x40;
	std::vector<uint8_t>& pac = G.seq.pac;
	pac.assign(F.n / 4 + 16, 0);          // never more bases than characters
	int64_t n = 0, run_start = -1;
	const char* p = F.p;
	const char* end = F.p + F.n;
	while (p < end) {
		if (*p == '>') {
			const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
			const char* e = nl ? nl : end;
			const char* k = p + 1;
			while (k < e && *k != ' ' && *k != '\t') ++k;
			if (!G.chr.empty()) G.chr.back().size = n - G.chr.back().start;
			Chr c; c.start = n; c.name.assign(p + 1, (size_t)(k - p - 1));
			G.chr.push_back(c);
			p = e;
			continue;
		}
		const char* gt = (const char*)memchr(p, '>', (size_t)(end - p));
		const char* stop = gt ? gt : end;
		while (p < stop) {
			// fast path: four ACGT letters inside a run, byte aligned
			if ((n & 3) == 0 && run_start >= 0) {
				uint8_t* out = pac.data() + (n >> 2);
				const char* q = p;
				while (q + 4 <= stop) {
					const uint8_t a = T[(unsigned char)q[0]], b = T[(unsigned char)q[1]], c = T[(unsigned char)q[2]], d = T[(unsigned char)q[3]];
					if ((a | b | c | d) & 0xC0) break;
					*out++ = (uint8_t)(a << 6 | b << 4 | c << 2 | d);
					q += 4;
				}
				n += q - p;
				p = q;
				if (p >= stop) break;
			}
			const uint8_t v = T[(unsigned char)*p++];
			if (v & 0x40) continue;
			if (v & 0x80) { if (run_start >= 0) { G.runs.push_back(run_start); G.runs.push_back(n - run_start); run_start = -1; } }
			else if (run_start < 0) run_start = n;
			pac[(size_t)(n >> 2)] |= (uint8_t)((v & 3u) << (((~n) & 3) << 1));
			++n;
		}
	}
	pac.resize((size_t)((n + 3) / 4));
	G.seq.n = n;
	if (run_start >= 0) { G.runs.push_back(run_start); G.runs.push_back(n - run_start); }
	if (!G.chr.empty()) G.chr.back().size = n - G.chr.back().start;
	if (G.chr.empty()) { err = std::string("no sequence in ") + path; return false; }
	return true;
}

// All reads of a file.  A read whose sequence is one line of the (mapped) file is used where it lies; the letters of a
// read spread over several lines are gathered, line ends removed, in a side arena.
struct Reads
{
	std::vector<int32_t> name;           // the number the reference prints for the read
	std::vector<const char*> ptr;
	std::vector<int64_t> len;
	int64_t total = 0;                   // letters of all reads
	MappedFile file;
	std::deque<std::string> side;        // stable addresses
	size_t size() const { return name.size(); }
	const char* data(size_t i) const { return ptr[i]; }
	int64_t length(size_t i) const { return len[i]; }
	void add(int32_t nm, const char* p, int64_t n) { name.push_back(nm); ptr.push_back(p); len.push_back(n); total += n; }
};

// appends [b, e) to the arena without '\n' and '\r'
inline void append_letters(std::string& arena, const char* b, const char* e)
{
	while (b < e) {
		const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
		const char* stop = nl ? nl : e;
		if (memchr(b, '\r', (size_t)(stop - b))) { for (const char* q = b; q < stop; ++q) if (*q != '\r') arena.push_back(*q); }
		else arena.append(b, (size_t)(stop - b));
		b = nl ? nl + 1 : e;
	}
}

inline bool load_reads(const char* path, Reads& R, std::string& err)
{
	MappedFile& F = R.file;
	if (!F.open(path)) { err = std::string("cannot open ") + path; return false; }
	if (F.n == 0) return true;
	const char* p = F.p;
	const char* end = F.p + F.n;
	if (*p == '>') {
		// chang_fastqfile reads character by character: a '>' anywhere opens a header that runs to the end of its line,
		// everything else up to the next '>' is sequence
		int next = 0;
		while (p < end) {
			const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));      // p is at a '>'
			const char* seq = nl ? nl + 1 : end;
			// the common case: the sequence is the next line and the line after it is a header (or the end)
			const char* l1 = seq < end ? (const char*)memchr(seq, '\n', (size_t)(end - seq)) : NULL;
			const char* line_end = l1 ? l1 : end;
			const char* after = l1 ? l1 + 1 : end;
			if ((after == end || *after == '>') && !memchr(seq, '>', (size_t)(line_end - seq)) && !memchr(seq, '\r', (size_t)(line_end - seq))) {
				R.add(next++, seq, line_end - seq);
				p = after;
				continue;
			}
			const char* gt = seq < end ? (const char*)memchr(seq, '>', (size_t)(end - seq)) : NULL;
			const char* stop = gt ? gt : end;
			R.side.emplace_back();
			append_letters(R.side.back(), seq, stop);
			R.add(next++, R.side.back().data(), (int64_t)R.side.back().size());
			p = stop;
		}
	} else {
		// FASTQ: records of four lines, the second one is the sequence
		int next = 0;
		const char* line[4];
		const char* line_end[4];
		int k = 0;
		bool more = true;
		while (more && p < end) {
			const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
			line[k] = p; line_end[k] = nl ? nl : end;
			if (!nl) more = false; else p = nl + 1;
			if (++k == 4) {
				const char* b = line[1];
				const char* e = line_end[1];
				if (e > b && e[-1] == '\r') --e;
				R.add(++next, b, e - b);
				k = 0;
			}
		}
	}
	return true;
}

// Reads [first, first + count) packed as one volume.  A read made of upper-case ACGT only is packed once (its reverse
// strand is the reverse complement of the packed bases); any other read also gets its reverse strand packed explicitly
// behind the forward strands, built the way the reference builds it.
struct ReadBatch
{
	std::vector<uint8_t> pac;
	int64_t nbases = 0;
	std::vector<int32_t> offsz, len, fread, rread, rrc;
	std::vector<int64_t> bad;
	mecat_volume vol;
	mecat_ref_reads view()
	{
		vol.num_reads = (int32_t)(offsz.size() / 2); vol.num_bases = (int32_t)nbases; vol.start_read_id = 0;
		vol.offset_size = offsz.data(); vol.pac = pac.data();
		mecat_ref_reads r;
		r.num_reads = (int32_t)len.size(); r.vol = &vol; r.read_len = len.data(); r.fwd_read = fread.data(); r.rev_read = rread.data();
		r.rev_is_rc = rrc.data(); r.num_bad = (int64_t)bad.size(); r.bad = bad.data();
		return r;
	}

	// bits 0-1: code (anything but ACGT in either case packs as A); bit 7: not upper-case ACGT
	struct PackTable
	{
		uint8_t t[256];
		PackTable() { for (int c = 0; c < 256; ++c) { const int k = code_ci((unsigned char)c); t[c] = (uint8_t)((k < 0 ? 0 : k) | (upper_acgt((unsigned char)c) ? 0 : 0x80)); } }
	};
	static const uint8_t* table()
	{
		static const PackTable T;       // initialised once, thread-safely: several threads pack
		return T.t;
	}
	// n letters at base offset `base`; bytes shared with a neighbouring read are OR-ed in atomically (several threads pack)
	static bool pack_letters(uint8_t* pac, int64_t base, const char* s, int64_t n, std::vector<int64_t>& bad)
	{
		const uint8_t* t = table();
		const unsigned char* u = (const unsigned char*)s;
		unsigned flags = 0;
		int64_t i = 0;
		auto one = [&](int64_t k) {
			const uint8_t v = t[u[k]];
			if (v & 0x80) { flags |= 0x80; bad.push_back(base + k); }
			__atomic_fetch_or(&pac[(base + k) >> 2], (uint8_t)((v & 3) << (((~(base + k)) & 3) << 1)), __ATOMIC_RELAXED);
		};
		while (i < n && ((base + i) & 3)) one(i++);
		uint8_t* out = pac + ((base + i) >> 2);
		for (; i + 4 <= n; i += 4) {
			const uint8_t a = t[u[i]], b = t[u[i + 1]], c = t[u[i + 2]], d = t[u[i + 3]];
			if ((a | b | c | d) & 0x80) {
				flags |= 0x80;
				if (a & 0x80) bad.push_back(base + i);
				if (b & 0x80) bad.push_back(base + i + 1);
				if (c & 0x80) bad.push_back(base + i + 2);
				if (d & 0x80) bad.push_back(base + i + 3);
			}
			*out++ = (uint8_t)((a & 3) << 6 | (b & 3) << 4 | (c & 3) << 2 | (d & 3));
		}
		while (i < n) one(i++);
		return flags == 0;
	}

	void build(const Reads& R, size_t first, size_t count, int threads)
	{
		len.resize(count); fread.resize(count); rread.resize(count); rrc.resize(count);
		offsz.resize(2 * count);
		int64_t at = 0;
		for (size_t i = 0; i < count; ++i) {
			const int64_t n = R.length(first + i);
			len[i] = (int32_t)n; fread[i] = (int32_t)i; rread[i] = (int32_t)i; rrc[i] = 1;
			offsz[2 * i] = (int32_t)at; offsz[2 * i + 1] = (int32_t)n;
			at += n + 1;                      // one pad base between reads, like the reference's volumes
		}
		pac.assign((size_t)((at + 3) / 4 + 8), 0);
		if (threads < 1) threads = 1;
		if ((size_t)threads > count) threads = count ? (int)count : 1;
		std::vector<std::vector<int64_t>> tbad((size_t)threads);
		std::vector<uint8_t> plain(count, 1);
		auto first_at = [&](int64_t pos) {      // first read whose offset is >= pos
			size_t a = 0, b = count;
			while (a < b) { const size_t m = (a + b) / 2; if (offsz[2 * m] < pos) a = m + 1; else b = m; }
			return a;
		};
		auto work = [&](int t) {                // a contiguous range of reads with about 1/threads of the bases
			const size_t r0 = first_at(at * t / threads), r1 = first_at(at * (t + 1) / threads);
			for (size_t r = r0; r < r1; ++r) plain[r] = pack_letters(pac.data(), offsz[2 * r], R.data(first + r), len[r], tbad[(size_t)t]) ? 1 : 0;
		};
		if (threads == 1) work(0);
		else {
			std::vector<std::thread> pool;
			for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
			for (auto& th : pool) th.join();
		}
		bad.clear();
		for (auto& v : tbad) bad.insert(bad.end(), v.begin(), v.end());
		// explicit reverse strands of the reads that are not plain
		for (size_t r = 0; r < count; ++r) {
			if (plain[r]) continue;
			std::string s(R.data(first + r), (size_t)len[r]);
			std::reverse(s.begin(), s.end());
			for (char& c : s) c = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
			rread[r] = (int32_t)(offsz.size() / 2); rrc[r] = 0;
			offsz.push_back((int32_t)at); offsz.push_back(len[r]);
			pac.resize((size_t)((at + len[r] + 1 + 3) / 4 + 8), 0);
			pack_letters(pac.data(), at, s.data(), len[r], bad);
			at += len[r] + 1;
		}
		nbases = at;
	}
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

// The header of a SAM file (print_sam_header / print_sam_references / print_sam_program, src/mecat2ref/output.cpp:92-116)
inline void sam_header(std::string& out, const Genome& G, int argc, char* const argv[])
{
	char line[1400];
	out += "@HD\tVN:1.4\tSO:unknown\tGO:query\n";
	for (const Chr& c : G.chr) { snprintf(line, sizeof line, "@SQ\tSN:%s\tLN:%ld\n", c.name.c_str(), (long)c.size); out += line; }
	out += "@PG\tID:0\tVN:0.0.1\tCL:";
	for (int i = 0; i < argc; ++i) { out += argv[i]; out += ' '; }
	out += "\tPN:mecat2ref\n";
}

// output_cigar, output.cpp:118-155: hard clips for the unaligned read ends, runs of D (gap in the read), I (gap in the
// reference) and M
inline void sam_cigar(std::string& out, int qstart, int qend, int qsize, const char* qmap, const char* smap, int n)
{
	char tmp[32];
	auto op = [&](int len, char c) { snprintf(tmp, sizeof tmp, "%d%c", len, c); out += tmp; };
	if (qstart) op(qstart, 'H');
	int i = 0;
	while (i < n) {
		int j = i + 1;
		if (qmap[i] == '-') { while (j < n && qmap[j] == '-') ++j; op(j - i, 'D'); }
		else if (smap[i] == '-') { while (j < n && smap[j] == '-') ++j; op(j - i, 'I'); }
		else { while (j < n && qmap[j] != '-' && smap[j] != '-') ++j; op(j - i, 'M'); }
		i = j;
	}
	if (qend != qsize) op(qsize - qend, 'H');
}

// format 0 = ref (header line + the two alignment strings), 1 = m4, 2 = sam records.  `names`: the printed number of read r.
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
		} else if (format == 1) {
			double ident = (double)r.matches;        // print_m4_result counts the equal columns of the two strings
			ident = ident / (double)r.columns;
			ident *= 100.0;
			snprintf(line, sizeof line, "%d\t%s\t%.4f\t%d\t%d\t%d\t%d\t%d\t0\t%ld\t%ld\t%ld\n", id, c.name.c_str(), ident, r.vscore, r.dir ? 1 : 0, qb, qe,
			         r.qs, (long)(r.sb - c.start), (long)(r.se - c.start), (long)c.size);
			out += line;
		} else {
			// output_sam, output.cpp:157-191: coordinates of the strand that aligned, SEQ = the aligned read bases
			const char* qm = qstr + r.str_offset;
			snprintf(line, sizeof line, "%d\t%d\t%s\t%ld\t255\t", id, r.dir ? 0x10 : 0, c.name.c_str(), (long)(r.sb - c.start) + 1);
			out += line;
			sam_cigar(out, r.qb, r.qe, r.qs, qm, sstr + r.str_offset, r.columns);
			out += "\t*\t0\t0\t";
			for (int k = 0; k < r.columns; ++k) if (qm[k] != '-') out += qm[k];
			out += "\t*\n";
		}
	}
}

}  // namespace refio
