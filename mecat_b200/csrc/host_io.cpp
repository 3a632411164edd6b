// mecat_b200/csrc/host_io.cpp -- host-side data formats of the path (no device code).
//
// Keeps the reference's on-disk contract byte for byte:
//   split_raw_dataset / dump_volume / load_volume   src/common/split_database.cpp:222-266,136-181
//   FastaReader::read_one_seq                       src/common/fasta_reader.cpp:6-60
//   add_one_seq + PackedDB::set_char                src/common/split_database.cpp:104-119, packed_db.h:98-101
//   fileindex.txt                                   src/common/split_database.cpp:195-200,374-393
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <string>
#include <thread>
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
	// p[0..n): nucleotide letters; acgt: every letter is one of ACGTacgt (the common case, packed 8 letters at a time)
	void add(const char* seq, size_t n, bool acgt)
	{
		offsz.push_back((int32_t)curr);
		offsz.push_back((int32_t)n);
		const size_t need = (size_t)((curr + (int64_t)n + 1 + 3) / 4) + 9;
		if (pac.size() < need) pac.resize(need + need / 2, 0);
		const unsigned char* p = (const unsigned char*)seq;
		size_t i = 0;
		// same OR as PackedDB::set_char: codes > 3 spill into neighbours exactly like the reference
		for (; i < n && (curr & 3); ++i, ++curr) pac[curr >> 2] |= (uint8_t)(kEnc.t[p[i]] << (((~curr) & 3) << 1));
		if (acgt) {
			// A 0x41, C 0x43, G 0x47, T 0x54 (either case): bits 2..1 give 0, 1, 3, 2; x ^ (x >> 1) turns that into 0, 1, 2, 3.
			// Four codes in the bytes of a 32-bit word collapse into one output byte (first letter in the top bits) with
			// one multiply: the wanted terms land in bits 24..31, every cross term stays below bit 24.
			// (locals only inside the loop: a uint8_t store may alias any member, which would force `curr` through memory)
			uint8_t* __restrict out = pac.data() + (curr >> 2);
			const size_t groups = (n - i) / 8;
			const unsigned char* __restrict in = p + i;
			for (size_t g = 0; g < groups; ++g) {
				uint64_t x;
				memcpy(&x, in + 8 * g, 8);
				uint64_t c = (x >> 1) & 0x0303030303030303ull;
				c ^= (c >> 1) & 0x0101010101010101ull;
				const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
				out[2 * g] = (uint8_t)((lo * 0x40100401u) >> 24);
				out[2 * g + 1] = (uint8_t)((hi * 0x40100401u) >> 24);
			}
			i += 8 * groups; curr += (int64_t)(8 * groups);
		}
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
// One deliberate difference: the reference's BufferLineReader reports an EMPTY line as end of input once its last
// 16 MB buffer has been loaded (buffer_line_iterator.cpp:31,44), i.e. a blank line silently truncates a small file;
// here blank lines are skipped wherever they occur, which is what read_one_seq itself intends (fasta_reader.cpp:13).
// The file is mapped (or, when it cannot be, read) whole, so lines are stable (pointer, length) views and a record
// that is one clean line -- the usual long-read FASTA -- is packed straight from the mapping without a copy.
struct FastaStream
{
	const char* base = NULL;
	size_t size = 0, pos = 0;
	void* map = NULL;
	std::vector<char> owned;    // fallback when the input cannot be mapped
	bool ok = false;
	const char* held = NULL;    // one line of push-back
	size_t held_n = 0;
	bool saw_fastq = false;     // a '@' or '+' line was seen: record boundaries are not those of plain FASTA
	// a view of [lo, hi) of another stream's bytes (split_parallel: one per thread)
	FastaStream(const FastaStream& whole, size_t lo, size_t hi) : base(whole.base), size(hi), pos(lo), ok(true) {}
	explicit FastaStream(const char* path)
	{
		const int fd = open(path, O_RDONLY);
		if (fd < 0) return;
		struct stat st;
		if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
			void* m = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
			if (m != MAP_FAILED) {
				map = m; base = (const char*)m; size = (size_t)st.st_size;
				madvise(m, size, MADV_SEQUENTIAL);
			}
		}
		if (!map) {                     // pipes, empty files, exotic file systems
			char tmp[1 << 16];
			ssize_t r;
			while ((r = read(fd, tmp, sizeof tmp)) > 0) owned.insert(owned.end(), tmp, tmp + r);
			base = owned.data(); size = owned.size();
		}
		close(fd);
		ok = true;
	}
	~FastaStream() { if (map) munmap(map, size); }
	// next line without its terminator ('\n', '\r\n' or '\r'); false at end of input
	bool line(const char*& p, size_t& n)
	{
		if (held) { p = held; n = held_n; held = NULL; return true; }
		if (pos >= size) return false;
		const char* s = base + pos;
		const size_t len = size - pos;
		const char* nl = (const char*)memchr(s, '\n', len);
		size_t i = nl ? (size_t)(nl - s) : len;
		const char* cr = (const char*)memchr(s, '\r', i);   // a lone '\r' also ends a line
		if (cr) i = (size_t)(cr - s);
		p = s; n = i;
		pos += i < len ? i + 1 : i;
		if (cr && pos < size && base[pos] == '\n') ++pos;   // the '\n' of a '\r\n' pair
		return true;
	}
	void unget(const char* p, size_t n) { held = p; held_n = n; }
	// Returns -1 at end of input, -2 on malformed input, else the sequence length; the sequence is [sp, sp + length)
	// (a view into the input or into `seq`), acgt tells whether every letter is one of ACGTacgt.
	int64_t next(std::string& seq, const char*& sp, bool& acgt, std::string& err)
	{
		seq.clear();
		sp = NULL; acgt = true;
		size_t view_n = 0;
		bool need_defline = true, got_defline = false;
		const char* l;
		size_t n;
		while (line(l, n)) {
			if (n == 0) continue;
			const int c = (unsigned char)l[0];
			if (c == '@' || c == '+') saw_fastq = true;
			if (c == '>' || c == '@') {
				if (need_defline) { need_defline = false; got_defline = true; continue; }
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
			// fast path: a line made only of nucleotide letters is taken as is
			unsigned bits = 0;
			for (size_t q = 0; q < n; ++q) bits |= kEnc.t[(unsigned char)l[q]];   // 16 only for non-nucleotides, 4 | 8 for non-ACGT codes
			if (bits & 12u) acgt = false;
			if (!(bits & 16u)) {
				if (!sp && seq.empty()) { sp = l; view_n = n; continue; }          // first data line: keep the view
				if (sp) { seq.assign(sp, view_n); sp = NULL; }                      // a second line: fall back to a copy
				seq.append(l, n);
				continue;
			}
			if (sp) { seq.assign(sp, view_n); sp = NULL; }
			for (size_t p = 0; p < n; ++p) {
				const int ch = (unsigned char)l[p];
				if (ch == ';') break;
				if (kEnc.t[ch] < 16) { seq.push_back((char)ch); if (kEnc.t[ch] > 3) acgt = false; }
				else if (!(ch == ' ' || (ch >= 9 && ch <= 13))) { err = "invalid residue in sequence data"; return -2; }
			}
		}
		if (sp) return (int64_t)view_n;
		if (seq.empty() && got_defline) { err = "sequence data is missing"; return -2; }
		if (!got_defline && seq.empty()) return -1;
		sp = seq.data();
		return (int64_t)seq.size();
	}
};

std::string join(const char* dir, const std::string& name)
{
	std::string p(dir);
	if (p.empty() || p[p.size() - 1] != '/') p += '/';
	return p + name;
}

// ---- the split on several host threads (plain FASTA only; anything else takes the sequential path above)
//
// The sequential split parses and packs ~1.1 GB of FASTA per second -- 13 s for the 15 GB of BASELINE configs[4], more than
// the 36 tiles take on 8 GPUs.  In plain FASTA a '>' at the start of a line always begins a record, so the file can be cut
// at such lines and parsed by one thread per piece (same record rules: FastaStream::next); the volume a read lands in
// depends on all reads before it, which is a cheap sequential pass over (length) records; packing is then parallel again,
// volume by volume, with the first and last partial byte of a read OR-ed in atomically (neighbouring reads share bytes),
// and a finished volume is written while the next one is packed.  A '@' or '+' line anywhere (FASTQ), a malformed record
// or an input that cannot be mapped makes the whole call fall back to the sequential path, which also owns the error texts.
struct ParsedRead { const char* sp; uint32_t n; bool acgt; };

inline void or_byte(uint8_t* p, uint8_t v) { __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

// VolumeBuilder::add into a shared, zeroed buffer at base offset `curr`
void pack_read_at(uint8_t* pac, int64_t curr, const char* seq, size_t n, bool acgt)
{
	const unsigned char* p = (const unsigned char*)seq;
	size_t i = 0;
	for (; i < n && (curr & 3); ++i, ++curr) or_byte(pac + (curr >> 2), (uint8_t)(kEnc.t[p[i]] << (((~curr) & 3) << 1)));
	if (acgt) {
		uint8_t* __restrict out = pac + (curr >> 2);
		const size_t groups = (n - i) / 8;
		const unsigned char* __restrict in = p + i;
		for (size_t g = 0; g < groups; ++g) {
			uint64_t x;
			memcpy(&x, in + 8 * g, 8);
			uint64_t c = (x >> 1) & 0x0303030303030303ull;
			c ^= (c >> 1) & 0x0101010101010101ull;
			const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
			out[2 * g] = (uint8_t)((lo * 0x40100401u) >> 24);
			out[2 * g + 1] = (uint8_t)((hi * 0x40100401u) >> 24);
		}
		i += 8 * groups; curr += (int64_t)(8 * groups);
	}
	// whole bytes belong to this read alone; a code > 3 may spill into the byte's other bases exactly like PackedDB::set_char
	for (; i + 4 <= n; i += 4, curr += 4) {
		const unsigned a = kEnc.t[p[i]], b = kEnc.t[p[i + 1]], c = kEnc.t[p[i + 2]], d = kEnc.t[p[i + 3]];
		pac[curr >> 2] = (uint8_t)((a << 6) | (b << 4) | (c << 2) | d);
	}
	for (; i < n; ++i, ++curr) or_byte(pac + (curr >> 2), (uint8_t)(kEnc.t[p[i]] << (((~curr) & 3) << 1)));
}

int split_threads(size_t input_bytes)
{
	if (const char* e = getenv("MECAT_B200_SPLIT_THREADS")) return std::max(1, atoi(e));     // 1 = sequential path; test hook for small files
	if (input_bytes < (64u << 20)) return 1;
	const int hw = (int)std::thread::hardware_concurrency();
	return std::max(1, std::min(hw > 0 ? hw : 1, 32));
}

// What becomes of a finished volume: a file of the work directory (mecat2pw) or a volume in host memory (mecat2cns).
// Called in volume order from one thread at a time; the buffers stay valid until the call returns.
typedef std::function<int(int vi, int num_reads, int64_t num_bases, int start_read_id, const std::vector<int32_t>& offsz,
                          const uint8_t* pac, size_t bytes)> VolumeSink;

// returns 0 = done, 1 = the sink failed (err set), 2 = not applicable: use the sequential path
int split_parallel(const FastaStream& in, int64_t cap, int threads, int* num_volumes, std::string& err, const VolumeSink& sink)
{
	if (!in.map || threads < 2 || in.size == 0) return 2;
	{
		// plain FASTA starts (after blank and comment lines) with '>'
		size_t q = 0;
		while (q < in.size) {
			const char c = in.base[q];
			if (c == '\n' || c == '\r') { ++q; continue; }
			if (c == '#' || c == '!') { while (q < in.size && in.base[q] != '\n' && in.base[q] != '\r') ++q; continue; }
			break;
		}
		if (q >= in.size || in.base[q] != '>') return 2;
	}
	// cut points: the first line start holding '>' at or after every nominal boundary
	std::vector<size_t> cut((size_t)threads + 1, in.size);
	cut[0] = 0;
	for (int t = 1; t < threads; ++t) {
		size_t q = in.size / (size_t)threads * (size_t)t;
		if (q < cut[(size_t)t - 1]) q = cut[(size_t)t - 1];
		for (;;) {
			while (q < in.size && in.base[q] != '\n' && in.base[q] != '\r') ++q;
			while (q < in.size && (in.base[q] == '\n' || in.base[q] == '\r')) ++q;
			if (q >= in.size || in.base[q] == '>') break;
		}
		cut[(size_t)t] = std::min(q, in.size);
	}
	const bool timing = getenv("MECAT_B200_SPLIT_TIMING") != NULL;
	auto now = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
	const double t_start = now();
	struct Piece { std::vector<ParsedRead> reads; std::deque<std::string> owned; bool bad = false; };
	std::vector<Piece> pieces((size_t)threads);
	{
		std::vector<std::thread> pool;
		for (int t = 0; t < threads; ++t)
			pool.emplace_back([&, t]() {
				Piece& P = pieces[(size_t)t];
				if (cut[(size_t)t] >= cut[(size_t)t + 1]) return;
				FastaStream part(in, cut[(size_t)t], cut[(size_t)t + 1]);
				std::string seq, e;
				const char* sp = NULL;
				bool acgt = true;
				for (;;) {
					const int64_t n = part.next(seq, sp, acgt, e);
					if (n == -1) break;
					if (n == -2 || part.saw_fastq || n > 0x7fffffffLL) { P.bad = true; return; }
					ParsedRead r;
					if (sp == seq.data()) { P.owned.push_back(seq); sp = P.owned.back().data(); }    // a copy was needed: keep it
					r.sp = sp; r.n = (uint32_t)n; r.acgt = acgt;
					P.reads.push_back(r);
				}
				if (part.saw_fastq) P.bad = true;
			});
		for (auto& th : pool) th.join();
	}
	for (const Piece& P : pieces) if (P.bad) return 2;
	const double t_parsed = now();
	// volumes: the reference's rule, read by read (split_database.cpp:240-259)
	struct Vol { size_t first_piece, first_read; int64_t curr = 0; int num_reads = 0; std::vector<int32_t> offsz; };
	std::vector<Vol> vols;
	std::vector<std::vector<std::pair<int, int64_t>>> where((size_t)threads);     // per read: volume, base offset
	{
		Vol v; v.first_piece = 0; v.first_read = 0;
		for (int t = 0; t < threads; ++t) {
			where[(size_t)t].resize(pieces[(size_t)t].reads.size());
			for (size_t i = 0; i < pieces[(size_t)t].reads.size(); ++i) {
				const int64_t n = pieces[(size_t)t].reads[i].n;
				if (v.curr + n + 1 > cap && v.curr > 0) { vols.push_back(std::move(v)); v = Vol(); }
				if (n + 1 > cap) return 2;                       // the sequential path reports it
				where[(size_t)t][i] = std::make_pair((int)vols.size(), v.curr);
				v.offsz.push_back((int32_t)v.curr); v.offsz.push_back((int32_t)n);
				v.curr += n + 1; ++v.num_reads;
			}
		}
		if (v.curr > 0) vols.push_back(std::move(v));
	}
	std::future<int> writer;
	int rid = 0, rc = 0;
	for (size_t vi = 0; vi < vols.size() && !rc; ++vi) {
		const Vol& V = vols[vi];
		const size_t bytes = (size_t)((V.curr + 3) / 4);
		std::shared_ptr<std::vector<uint8_t>> pac(new std::vector<uint8_t>(bytes + 16, 0));
		{
			std::atomic<size_t> next(0);
			const size_t grain = 256;
			// flat list of (piece, read) of this volume, handed out in grains
			std::vector<std::pair<int, size_t>> items;
			items.reserve((size_t)V.num_reads);
			for (int t = 0; t < threads; ++t)
				for (size_t i = 0; i < where[(size_t)t].size(); ++i)
					if (where[(size_t)t][i].first == (int)vi) items.push_back(std::make_pair(t, i));
			std::vector<std::thread> pool;
			for (int t = 0; t < threads; ++t)
				pool.emplace_back([&]() {
					for (size_t a; (a = next.fetch_add(grain)) < items.size();)
						for (size_t k = a; k < std::min(items.size(), a + grain); ++k) {
							const ParsedRead& r = pieces[(size_t)items[k].first].reads[items[k].second];
							pack_read_at(pac->data(), where[(size_t)items[k].first][items[k].second].second, r.sp, r.n, r.acgt);
						}
				});
			for (auto& th : pool) th.join();
		}
		if (writer.valid() && writer.get()) { rc = 1; break; }
		const int first_id = rid;
		rid += V.num_reads;
		const Vol* vp = &V;
		const int vnum = (int)vi;
		// the sink of volume vi runs behind the packing of volume vi + 1
		writer = std::async(std::launch::async, [&sink, vp, vnum, first_id, pac, bytes]() -> int {
			return sink(vnum, vp->num_reads, vp->curr, first_id, vp->offsz, pac->data(), bytes);
		});
	}
	const double t_packed = now();
	if (writer.valid() && writer.get()) rc = 1;
	if (rc) { err = "cannot store a volume"; return 1; }
	*num_volumes = (int)vols.size();
	if (timing) fprintf(stderr, "[split] %d threads: parse %.2f s, assign + pack (+ writes behind it) %.2f s, last write %.2f s\n", threads,
	                    t_parsed - t_start, t_packed - t_parsed, now() - t_packed);
	return 0;
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
	if (!in.ok) return fail(std::string("cannot open file '") + reads_path + "' for reading");
	{
		// volume files + fileindex.txt, written as the volumes come
		std::string perr;
		std::vector<std::string> names;
		const VolumeSink to_file = [&](int vi, int num_reads, int64_t num_bases, int start_read_id, const std::vector<int32_t>& offsz,
		                               const uint8_t* pac, size_t bytes) -> int {
			const std::string name = join(wrk_dir, "vol" + std::to_string(vi));
			names.push_back(name);
			FILE* f = fopen(name.c_str(), "wb");
			if (!f) return 1;
			const int32_t hdr[3] = {num_reads, (int32_t)num_bases, start_read_id};
			bool ok = fwrite(hdr, 4, 3, f) == 3 && fwrite(offsz.data(), 4, offsz.size(), f) == offsz.size() && fwrite(pac, 1, bytes, f) == bytes;
			ok = (fclose(f) == 0) && ok;
			return ok ? 0 : 1;
		};
		const int prc = split_parallel(in, cap, split_threads(in.size), num_volumes, perr, to_file);
		if (prc == 0) {
			FILE* idx = fopen(join(wrk_dir, "fileindex.txt").c_str(), "w");
			if (!idx) return fail(std::string("cannot write into '") + wrk_dir + "'");
			for (const std::string& n : names) fprintf(idx, "%s\n", n.c_str());
			fclose(idx);
			return 0;
		}
		if (prc == 1) return fail("cannot write volume file");
	}
	FILE* idx = fopen(join(wrk_dir, "fileindex.txt").c_str(), "w");
	if (!idx) return fail(std::string("cannot write into '") + wrk_dir + "'");
	VolumeBuilder v;
	{
		// one allocation up front: a volume never holds more bases than the input has bytes (growing a vector of this size
		// by reallocation cost more than the packing itself)
		const int64_t bases = std::min<int64_t>(cap, (int64_t)in.size) + 1;
		v.pac.assign((size_t)(bases / 4) + 64, 0);
		v.offsz.reserve((size_t)std::min<int64_t>(bases / 1000 + 16, 1 << 24));
	}
	int vol = 0, rid = 0;
	std::string seq, e;
	const char* sp = NULL;
	bool acgt = true;
	auto flush = [&]() -> int {
		const std::string name = join(wrk_dir, "vol" + std::to_string(vol++));
		fprintf(idx, "%s\n", name.c_str());
		if (v.dump(name.c_str(), rid)) return 1;
		rid += v.num_reads;
		v.clear();
		return 0;
	};
	for (;;) {
		const int64_t n = in.next(seq, sp, acgt, e);
		if (n == -1) break;
		if (n == -2) { fclose(idx); return fail("FastaReader: " + e); }
		if (v.curr + n + 1 > cap && v.curr > 0) { if (flush()) { fclose(idx); return fail("cannot write volume file"); } }
		if (n + 1 > cap) { fclose(idx); return fail("a read is longer than the volume cap"); }
		v.add(sp, (size_t)n, acgt);
	}
	if (v.curr > 0 && flush()) { fclose(idx); return fail("cannot write volume file"); }
	fclose(idx);
	*num_volumes = vol;
	return 0;
}

// load_volume (split_database.cpp:156-181).  Buffers are malloc'ed; release with
// mecat_b200_volume_unload.
// All reads of a FASTA/FASTQ file as ONE packed volume in memory (PackedDB::load_fasta_db, src/common/packed_db.cpp:194,
// as mecat2cns uses it): the same records and bytes split_dataset would write into vol0, without the file round trip.
int mecat_b200_volume_from_fasta(const char* reads_path, mecat_volume* out, char* err, int err_cap)
{
	auto fail = [&](const std::string& m) { if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", m.c_str()); return 1; };
	if (!reads_path || !out) return fail("volume_from_fasta: null argument");
	FastaStream in(reads_path);
	if (!in.ok) return fail(std::string("cannot open file '") + reads_path + "' for reading");
	VolumeBuilder v;
	const int64_t bases = std::min<int64_t>(kMaxVolumeBases, (int64_t)in.size) + 1;
	v.pac.assign((size_t)(bases / 4) + 64, 0);
	std::string seq, e;
	const char* sp = NULL;
	bool acgt = true;
	for (;;) {
		const int64_t n = in.next(seq, sp, acgt, e);
		if (n == -1) break;
		if (n == -2) return fail("FastaReader: " + e);
		if (v.curr + n + 1 > kMaxVolumeBases) return fail("the read set needs more than one 2.14 Gbase volume");
		v.add(sp, (size_t)n, acgt);
	}
	const size_t nr = (size_t)v.num_reads, bytes = (size_t)((v.curr + 3) / 4);
	int32_t* os = (int32_t*)malloc(sizeof(int32_t) * 2 * (nr ? nr : 1));
	uint8_t* pac = (uint8_t*)malloc(bytes + 16);
	if (!os || !pac) { free(os); free(pac); return fail("out of memory"); }
	if (nr) memcpy(os, v.offsz.data(), sizeof(int32_t) * 2 * nr);
	memcpy(pac, v.pac.data(), bytes);
	memset(pac + bytes, 0, 16);
	out->num_reads = v.num_reads; out->num_bases = (int32_t)v.curr; out->start_read_id = 0;
	out->offset_size = os; out->pac = pac;
	return 0;
}

// The whole read set as volumes in host memory (several when it exceeds the cap): what mecat2cns needs for read sets
// larger than 2.14 Gbase.  Same bytes as the files mecat_b200_split_dataset writes.  Release with mecat_b200_volumes_unload.
int mecat_b200_volumes_from_fasta(const char* reads_path, int64_t max_volume_bases, mecat_volume** vols_out, int* num_volumes,
                                  char* err, int err_cap)
{
	auto fail = [&](const std::string& m) { if (err && err_cap > 0) snprintf(err, (size_t)err_cap, "%s", m.c_str()); return 1; };
	if (!reads_path || !vols_out || !num_volumes) return fail("volumes_from_fasta: null argument");
	const int64_t cap = max_volume_bases > 0 ? max_volume_bases : kMaxVolumeBases;
	FastaStream in(reads_path);
	if (!in.ok) return fail(std::string("cannot open file '") + reads_path + "' for reading");
	std::vector<mecat_volume> got;
	bool oom = false;
	auto keep = [&](int num_reads, int64_t num_bases, int start_read_id, const int32_t* offsz, const uint8_t* pac, size_t bytes) -> int {
		const size_t nr = (size_t)num_reads;
		int32_t* os = (int32_t*)malloc(sizeof(int32_t) * 2 * (nr ? nr : 1));
		uint8_t* pc = (uint8_t*)malloc(bytes + 16);
		if (!os || !pc) { free(os); free(pc); oom = true; return 1; }
		if (nr) memcpy(os, offsz, sizeof(int32_t) * 2 * nr);
		memcpy(pc, pac, bytes);
		memset(pc + bytes, 0, 16);
		mecat_volume v;
		v.num_reads = num_reads; v.num_bases = (int32_t)num_bases; v.start_read_id = start_read_id; v.offset_size = os; v.pac = pc;
		got.push_back(v);
		return 0;
	};
	auto drop_all = [&]() { for (mecat_volume& v : got) mecat_b200_volume_unload(&v); got.clear(); };
	{
		std::string perr;
		int nv = 0;
		const VolumeSink to_memory = [&](int, int num_reads, int64_t num_bases, int start_read_id, const std::vector<int32_t>& offsz,
		                                 const uint8_t* pac, size_t bytes) -> int { return keep(num_reads, num_bases, start_read_id, offsz.data(), pac, bytes); };
		const int prc = split_parallel(in, cap, split_threads(in.size), &nv, perr, to_memory);
		if (prc == 1) { drop_all(); return fail(oom ? "out of memory" : perr); }
		if (prc == 2) {
			drop_all();
			VolumeBuilder v;
			const int64_t bases = std::min<int64_t>(cap, (int64_t)in.size) + 1;
			v.pac.assign((size_t)(bases / 4) + 64, 0);
			std::string seq, e;
			const char* sp = NULL;
			bool acgt = true;
			int rid = 0;
			auto flush = [&]() -> int {
				if (keep(v.num_reads, v.curr, rid, v.offsz.data(), v.pac.data(), (size_t)((v.curr + 3) / 4))) return 1;
				rid += v.num_reads;
				v.clear();
				return 0;
			};
			for (;;) {
				const int64_t n = in.next(seq, sp, acgt, e);
				if (n == -1) break;
				if (n == -2) { drop_all(); return fail("FastaReader: " + e); }
				if (v.curr + n + 1 > cap && v.curr > 0) { if (flush()) { drop_all(); return fail("out of memory"); } }
				if (n + 1 > cap) { drop_all(); return fail("a read is longer than the volume cap"); }
				v.add(sp, (size_t)n, acgt);
			}
			if (v.curr > 0 && flush()) { drop_all(); return fail("out of memory"); }
		}
	}
	mecat_volume* arr = (mecat_volume*)malloc(sizeof(mecat_volume) * (got.size() ? got.size() : 1));
	if (!arr) { drop_all(); return fail("out of memory"); }
	for (size_t i = 0; i < got.size(); ++i) arr[i] = got[i];
	*vols_out = arr; *num_volumes = (int)got.size();
	return 0;
}

void mecat_b200_volumes_unload(mecat_volume* vols, int num_volumes)
{
	if (!vols) return;
	for (int i = 0; i < num_volumes; ++i) mecat_b200_volume_unload(&vols[i]);
	free(vols);
}

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
