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
	void clear() { offsz.clear(); pac.clear(); curr = 0; num_reads = 0; }
	void add(const std::string& s)
	{
		offsz.push_back((int32_t)curr);
		offsz.push_back((int32_t)s.size());
		const size_t need = (size_t)((curr + (int64_t)s.size() + 1 + 3) / 4) + 1;
		if (pac.size() < need) pac.resize(need, 0);
		for (size_t i = 0; i < s.size(); ++i, ++curr) {
			const uint8_t c = kEnc.t[(unsigned char)s[i]];
			// same OR as PackedDB::set_char: codes > 3 spill into neighbours exactly like the reference
			pac[curr >> 2] |= (uint8_t)(c << (((~curr) & 3) << 1));
		}
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
struct FastaStream
{
	FILE* f;
	std::vector<char> buf;
	std::string pending;
	bool have_pending = false;
	explicit FastaStream(const char* path) : f(fopen(path, "rb")), buf(8u << 20) { if (f) setvbuf(f, buf.data(), _IOFBF, buf.size()); }
	~FastaStream() { if (f) fclose(f); }
	bool line(std::string& out)
	{
		if (have_pending) { out.swap(pending); have_pending = false; return true; }
		out.clear();
		int c;
		bool any = false;
		while ((c = getc_unlocked(f)) != EOF) {
			any = true;
			if (c == '\n') return true;
			if (c == '\r') { int d = getc_unlocked(f); if (d != '\n' && d != EOF) ungetc(d, f); return true; }
			out.push_back((char)c);
		}
		return any;
	}
	void unget(std::string& l) { pending.swap(l); have_pending = true; }
	// returns -1 at end of input, -2 on malformed input, else the sequence length
	int64_t next(std::string& seq, std::string& err)
	{
		seq.clear();
		bool need_defline = true, got_defline = false;
		std::string l;
		while (line(l)) {
			if (l.empty()) continue;
			const int c = (unsigned char)l[0];
			if (c == '>' || c == '@') {
				if (need_defline) { need_defline = false; got_defline = true; continue; }
				unget(l);
				break;
			} else if (c == '+') {
				std::string q;
				if (!line(q)) { err = "quality score line is missing"; return -2; }
				break;
			} else if (c == '#' || c == '!') {
				continue;
			} else if (need_defline) {
				err = "input doesn't start with a defline or comment";
				return -2;
			}
			for (size_t p = 0; p < l.size(); ++p) {
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
