// mecat_b200/csrc/host_io.cpp -- host-side data formats of the path (no device code).
//
// Keeps the reference's on-disk contract byte for byte:
//   split_raw_dataset / dump_volume / load_volume   src/common/split_database.cpp:222-266,136-181
//   FastaReader::read_one_seq                       src/common/fasta_reader.cpp:6-60
//   add_one_seq + PackedDB::set_char                src/common/split_database.cpp:104-119, packed_db.h:98-101
//   fileindex.txt                                   src/common/split_database.cpp:195-200,374-393
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mecat_b200.h"

namespace {

const int64_t kMaxVolumeBases = 2140000000LL;   // MCS, split_database.h:6

// get_dna_encode_table, src/common/defs.cpp:3-42 (IUPAC codes; 16 = not a nucleotide)
struct EncodeTable
{
	uint8_t t[256];
	EncodeTable()
	{
		memset(t, 16, sizeof t);
		const char* lo = "-acmgrsvtwyhkdbn";
		const uint8_t val[] = {15, 0, 1, 6, 2, 4, 9, 13, 3, 8, 5, 12, 7, 11, 10, 14};
		for (int i = 0; lo[i]; ++i) {
			t[(unsigned char)lo[i]] = val[i];
			if (lo[i] != '-') t[(unsigned char)(lo[i] - 'a' + 'A')] = val[i];
		}
	}
};
const EncodeTable kEnc;

struct VolumeBuilder
{
	std::vector<int32_t> offsz;
	std::vector<uint8_t> pac;
	int64_t curr = 0;
	int num_reads = 0;
	void clear() { offsz.clear(); pac.assign(pac.size(), 0); curr = 0; num_reads = 0; }
	void add(const std::string& s)
	{
		offsz.push_back((int32_t)curr);
		offsz.push_back((int32_t)s.size());
		const size_t need = (size_t)((curr + (int64_t)s.size() + 1 + 3) / 4) + 1;
		if (pac.size() < need) pac.resize(need + need / 2, 0);
		const unsigned char* p = (const unsigned char*)s.data();
		size_t n = s.size(), i = 0;
		// same OR as PackedDB::set_char: codes > 3 spill into neighbours exactly like the reference
		for (; i < n && (curr & 3); ++i, ++curr) pac[curr >> 2] |= (uint8_t)(kEnc.t[p[i]] << (((~curr) & 3) << 1));
		for (; i + 4 <= n; i += 4, curr += 4) {
			const unsigned a = kEnc.t[p[i]], b = kEnc.t[p[i + 1]], c = kEnc.t[p[i + 2]], d = kEnc.t[p[i + 3]];
			pac[curr >> 2] |= (uint8_t)((a << 6) | (b << 4) | (c << 2) | d);
		}
		for (; i < n; ++i, ++curr) pac[curr >> 2] |= (uint8_t)(kEnc.t[p[i]] << (((~curr) & 3) << 1));
		++curr;   // pad base, split_database.cpp:251
		++num_reads;
	}
	int dump(const char* path, int start_read_id) const
	{
		FILE* f = fopen(path, "wb");
		if (!f) return 1;
		const int32_t hdr[3] = {num_reads, (int32_t)curr, start_read_id};
		const size_t bytes = (size_t)((curr + 3) / 4);
		bool ok = fwrite(hdr, 4, 3, f) == 3 && fwrite(offsz.data(), 4, offsz.size(), f) == offsz.size() &&
		          fwrite(pac.data(), 1, bytes, f) == bytes;
		ok = (fclose(f) == 0) && ok;
		return ok ? 0 : 1;
	}
};

// Line reader with the reference's record rules: '>' or '@' starts a record, '+' ends it and
// swallows one quality line, '#'/'!' lines are comments, data lines stop at ';'.
// Block-buffered: lines are returned as (pointer, length) views into a 16 MB window.
struct FastaStream
{
	FILE* f;
	std::vector<char> buf;
	size_t beg = 0, end = 0;
	bool eof = false;
	std::string carry;          // a line that straddles two windows
	const char* held = NULL;    // one line of push-back
	size_t held_n = 0;
	explicit FastaStream(const char* path) : f(fopen(path, "rb")), buf(16u << 20) {}
	~FastaStream() { if (f) fclose(f); }
	bool fill()
	{
		if (eof) return false;
		beg = 0;
		end = fread(buf.data(), 1, buf.size(), f);
		if (end == 0) { eof = true; return false; }
		return true;
	}
	// next line without its terminator ('\n', '\r\n' or '\r'); false at end of input
	bool line(const char*& p, size_t& n)
	{
		if (held) { p = held; n = held_n; held = NULL; return true; }
		carry.clear();
		bool any = false;
		for (;;) {
			if (beg == end && !fill()) {
				if (!any) return false;
				p = carry.data(); n = carry.size();
				return true;
			}
			any = true;
			const char* s = buf.data() + beg;
			size_t len = end - beg, i;
			{
				const char* nl = (const char*)memchr(s, '\n', len);
				i = nl ? (size_t)(nl - s) : len;
				const char* cr = (const char*)memchr(s, '\r', i);   // a lone '\r' also ends a line
				if (cr) i = (size_t)(cr - s);
			}
			if (i < len) {
				const bool cr = s[i] == '\r';
				if (carry.empty()) { p = s; n = i; }
				else { carry.append(s, i); p = carry.data(); n = carry.size(); }
				beg += i + 1;
				if (cr) {                                  // swallow the '\n' of a '\r\n' pair
					if (beg == end) { std::string keep(p, n); fill(); carry.swap(keep); p = carry.data(); n = carry.size(); }
					if (beg < end && buf[beg] == '\n') ++beg;
				}
				return true;
			}
			carry.append(s, len);
			beg = end;
		}
	}
	void unget(const char* p, size_t n) { held = p; held_n = n; }
	// returns -1 at end of input, -2 on malformed input, else the sequence length
	int64_t next(std::string& seq, std::string& err)
	{
		seq.clear();
		bool need_defline = true, got_defline = false;
		const char* l;
		size_t n;
		std::string keep;
		while (line(l, n)) {
			if (n == 0) continue;
			const int c = (unsigned char)l[0];
			if (c == '>' || c == '@') {
				if (need_defline) { need_defline = false; got_defline = true; continue; }
				if (l == carry.data()) { keep.assign(l, n); carry.swap(keep); l = carry.data(); }
				unget(l, n);
				break;
			} else if (c == '+') {
				const char* q; size_t qn;
				if (!line(q, qn)) { err = "quality score line is missing"; return -2; }
				break;
			} else if (c == '#' || c == '!') {
				continue;
			} else if (need_defline) {
				err = "input doesn't start with a defline or comment";
				return -2;
			}
			// fast path: a line made only of nucleotide letters is appended as is
			size_t p = 0;
			{
				unsigned bad = 0;
				for (size_t q = 0; q < n; ++q) bad |= kEnc.t[(unsigned char)l[q]];   // 16 only for non-nucleotides
				p = (bad & 16u) ? 0 : n;
			}
			if (p == n) { seq.append(l, n); continue; }
			for (p = 0; p < n; ++p) {
				const int ch = (unsigned char)l[p];
				if (ch == ';') break;
				if (kEnc.t[ch] < 16) seq.push_back((char)ch);
				else if (!(ch == ' ' || (ch >= 9 && ch <= 13))) { err = "invalid residue in sequence data"; return -2; }
			}
		}
		if (seq.empty() && got_defline) { err = "sequence data is missing"; return -2; }
		if (!got_defline && seq.empty()) return -1;
		return (int64_t)seq.size();
	}
};

std::string join(const char* dir, const std::string& name)
{
	std::string p(dir);
	if (p.empty() || p[p.size() - 1] != '/') p += '/';
	return p + name;
}

}  // namespace

extern "C" {

// split_raw_dataset (split_database.cpp:222-266).  max_volume_bases <= 0 selects the
// reference's MCS; a smaller cap exists for multi-volume tests (the reference's own
// commented-out debug value, split_database.h:7).
int mecat_b200_split_dataset(const char* reads_path, const char* wrk_dir, int64_t max_volume_bases, int* num_volumes,
                             char* err, int err_cap)
{
	auto fail = [&](const std::string& m) { if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", m.c_str()); return 1; };
	if (!reads_path || !wrk_dir || !num_volumes) return fail("split_dataset: null argument");
	const int64_t cap = max_volume_bases > 0 ? max_volume_bases : kMaxVolumeBases;
	FastaStream in(reads_path);
	if (!in.f) return fail(std::string("cannot open file '") + reads_path + "' for reading");
	FILE* idx = fopen(join(wrk_dir, "fileindex.txt").c_str(), "w");
	if (!idx) return fail(std::string("cannot write into '") + wrk_dir + "'");
	VolumeBuilder v;
	int vol = 0, rid = 0;
	std::string seq, e;
	auto flush = [&]() -> int {
		const std::string name = join(wrk_dir, "vol" + std::to_string(vol++));
		fprintf(idx, "%s\n", name.c_str());
		if (v.dump(name.c_str(), rid)) return 1;
		rid += v.num_reads;
		v.clear();
		return 0;
	};
	for (;;) {
		const int64_t n = in.next(seq, e);
		if (n == -1) break;
		if (n == -2) { fclose(idx); return fail("FastaReader: " + e); }
		if (v.curr + n + 1 > cap && v.curr > 0) { if (flush()) { fclose(idx); return fail("cannot write volume file"); } }
		if (n + 1 > cap) { fclose(idx); return fail("a read is longer than the volume cap"); }
		v.add(seq);
	}
	if (v.curr > 0 && flush()) { fclose(idx); return fail("cannot write volume file"); }
	fclose(idx);
	*num_volumes = vol;
	return 0;
}

// load_volume (split_database.cpp:156-181).  Buffers are malloc'ed; release with
// mecat_b200_volume_unload.
int mecat_b200_volume_load(const char* path, mecat_volume* out)
{
	if (!path || !out) return 1;
	FILE* f = fopen(path, "rb");
	if (!f) return 2;
	int32_t hdr[3];
	if (fread(hdr, 4, 3, f) != 3 || hdr[0] < 0 || hdr[1] < 0) { fclose(f); return 3; }
	const size_t n = (size_t)hdr[0], bytes = ((size_t)hdr[1] + 3) / 4;
	int32_t* os = (int32_t*)malloc(sizeof(int32_t) * 2 * (n ? n : 1));
	uint8_t* pac = (uint8_t*)malloc(bytes + 16);
	if (!os || !pac || fread(os, 8, n, f) != n || fread(pac, 1, bytes, f) != bytes) { free(os); free(pac); fclose(f); return 3; }
	memset(pac + bytes, 0, 16);
	fclose(f);
	out->num_reads = hdr[0]; out->num_bases = hdr[1]; out->start_read_id = hdr[2];
	out->offset_size = os; out->pac = pac;
	return 0;
}

void mecat_b200_volume_unload(mecat_volume* v)
{
	if (!v) return;
	free((void*)v->offset_size);
	free((void*)v->pac);
	v->offset_size = NULL; v->pac = NULL;
}

}  // extern "C"
