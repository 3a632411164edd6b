// tests/ref_host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the product's mecat2ref stage sequence (mecat_b200/csrc/ref_pipeline.h), kernel bodies (ref_core.cuh) and host
// I/O (host/refio.h) on the host: "device memory" is malloc, a kernel launch is a loop over the units, the k-mer index
// is a plain two-pass CSR build, and the gapped extension -- a separate, GPU-tested kernel of the product (align.cu) --
// is played by the oracle's orc_diff_go.  This lets the CPU test-suite check the statements the GPU executes against
// the golden output of the unmodified mecat2ref binary without a GPU.  Compiled by tests/util.py into
// tests/_build/libref_harness.so; never part of the product library.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../include/mecat_b200.h"
#include "../mecat_b200/csrc/host/refio.h"
#include "../mecat_b200/csrc/ref_pipeline.h"
#include "../oracle/oracle.h"

namespace {

int g_tech = 0;      // -x of the run: which oracle aligner plays the extension kernel (harness_set_tech)

std::vector<uint32_t> words_of(const uint8_t* pac, int64_t n)      // the device layout of volume.cu: base p at bits 2 (p % 16) of word p / 16
{
	std::vector<uint32_t> w((size_t)(n / 16 + 9), 0u);
	for (int64_t p = 0; p < n; ++p) {
		const uint32_t b = (pac[p >> 2] >> (((~p) & 3) << 1)) & 3u;
		w[(size_t)(p >> 4)] |= b << ((p & 15) << 1);
	}
	return w;
}

struct HostBackend
{
	std::vector<void*> owned;
	std::string err;
	// what the extension hook needs: both volumes, unpacked
	const uint32_t* rfwd = nullptr; const int32_t* roffsz = nullptr;
	const uint32_t* gfwd = nullptr;
	int64_t tasks_run = 0, batches = 0, rescue_units = 0, string_tasks = 0;

	template <class T> T* alloc(size_t n)
	{
		void* p = malloc((n ? n : 1) * sizeof(T));
		if (!p) { err = "out of memory"; return nullptr; }
		memset(p, 0xAB, (n ? n : 1) * sizeof(T));      // like device memory: never zero by luck
		owned.push_back(p);
		return (T*)p;
	}
	template <class T> bool upload(T* d, const T* h, size_t n) { if (n) memcpy(d, h, n * sizeof(T)); return true; }
	template <class T> bool download(T* h, const T* d, size_t n) { if (n) memcpy(h, d, n * sizeof(T)); return true; }
	bool fill(void* d, int byte, size_t bytes) { memset(d, byte, bytes); return true; }
	bool launch_seed_warp(int64_t n, const mbref::SeedWarpFn& f, int stage)
	{
		if (stage == mbref::ST_SEED) ++batches;
		mbref::WarpScratch W;
		memset(&W, 0xAB, sizeof W);
		for (int64_t i = 0; i < n; ++i) f(i, mbref::EmuLanes(), W);
		return true;
	}
	template <class F> bool launch(int64_t n, const F& f, int stage)
	{
		if (stage == mbref::ST_SEED) ++batches;
		if (stage == mbref::ST_RESCUE) rescue_units += n;
		for (int64_t i = 0; i < n; ++i) f(i);
		return true;
	}
	bool release(void* p)
	{
		for (size_t i = 0; i < owned.size(); ++i) if (owned[i] == p) { free(p); owned[i] = owned.back(); owned.pop_back(); return true; }
		err = "release of an unknown block";
		return false;
	}
	void note_hits(int64_t) {}
	void fail(const char* m) { err = m; }
	void end_batch() { for (void* p : owned) free(p); owned.clear(); }

	static int base(const uint32_t* w, int64_t p) { return (int)((w[p >> 4] >> ((p & 15) << 1)) & 3u); }
	bool align(const mecat_align_task* tasks, size_t n, bool want_strings, mecat_align_result* res, std::vector<char>& qs, std::vector<char>& ss)
	{
		qs.clear(); ss.clear();
		std::vector<char> q, t, qa, ta;
		for (size_t i = 0; i < n; ++i) {
			const mecat_align_task& k = tasks[i];
			const int64_t off = roffsz[2 * k.qread];
			const int len = roffsz[2 * k.qread + 1];
			q.resize((size_t)len);
			for (int j = 0; j < len; ++j) q[(size_t)j] = (char)(k.qstrand ? 3 - base(rfwd, off + len - 1 - j) : base(rfwd, off + j));
			t.resize((size_t)k.swin_len);
			for (int j = 0; j < k.swin_len; ++j) t[(size_t)j] = (char)base(gfwd, (int64_t)k.swin_off + j);
			const int cap = 2 * (len + k.swin_len) + 64;
			qa.resize((size_t)cap); ta.resize((size_t)cap);
			int32_t o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
			double ident = 0;
			if (want_strings) ++string_tasks;
			const int ok = (g_tech == 1 ? orc_xdrop_go : orc_diff_go)(q.data(), k.qstart, len, t.data(), k.sstart, k.swin_len, 1000, o, &ident, qa.data(), ta.data(), cap);
			mecat_align_result& r = res[i];
			memset(&r, 0, sizeof r);
			r.ok = ok; r.str_offset = -1;
			if (!ok) continue;
			r.qstart = o[1]; r.qend = o[2]; r.sstart = o[3]; r.send = o[4]; r.columns = o[5]; r.matches = o[6]; r.ident = ident;
			if (want_strings) {
				r.str_offset = (int64_t)qs.size();
				qs.insert(qs.end(), qa.data(), qa.data() + o[5] + 1);
				ss.insert(ss.end(), ta.data(), ta.data() + o[5] + 1);
				qs.back() = 0; ss.back() = 0;
			}
			++tasks_run;
		}
		return true;
	}
};

// CSR k-mer index in the product's layout (index.cu): codes A0 C1 G2 T3 with the first base most significant, 0-based
// starts ascending, lists of more than 128 starts dropped.  `runs`: the {offset, size} "reads" the index is built over --
// the chunks the product cuts the genome's ACGT runs into (mbref::index_chunks), in ascending order.
void build_index(const uint32_t* gfwd, const std::vector<int32_t>& runs, std::vector<uint32_t>& begin, std::vector<int32_t>& pos)
{
	const uint32_t ncodes = 1u << 26, mask = ncodes - 1;
	std::vector<uint32_t> cnt(ncodes, 0u);
	for (int pass = 0; pass < 2; ++pass) {
		for (size_t r = 0; r + 1 < runs.size(); r += 2) {
			uint32_t code = 0;
			for (int64_t j = 0; j < runs[r + 1]; ++j) {
				const int64_t p = (int64_t)runs[r] + j;
				code = ((code << 2) | ((gfwd[p >> 4] >> ((p & 15) << 1)) & 3u)) & mask;
				if (j < 12) continue;
				if (pass == 0) ++cnt[code];
				else if (cnt[code] != 0xffffffffu) pos[begin[code] + cnt[code]++] = (int32_t)(p - 12);
			}
		}
		if (pass == 0) {
			begin.assign((size_t)ncodes + 1, 0u);
			uint32_t total = 0;
			for (uint32_t c = 0; c < ncodes; ++c) {
				begin[c] = total;
				if (cnt[c] > 128) cnt[c] = 0xffffffffu; else { total += cnt[c]; cnt[c] = 0; }
			}
			begin[ncodes] = total;
			pos.assign(total, 0);
		}
	}
}

// one ABI-shaped call on the host: what mecat_b200_ref_index_build + mecat_b200_ref_map do on the device
struct HostIndex { std::vector<uint32_t> gw, begin; std::vector<int32_t> pos; int64_t n = 0; };

void index_genome(const mecat_ref_genome* g, HostIndex& I)
{
	I.n = g->num_bases;
	I.gw = words_of(g->pac, g->num_bases);
	std::vector<int32_t> chunks;
	if (!mbref::index_chunks(g->run_start_len, g->num_runs, g->num_bases, chunks)) chunks.clear();
	build_index(I.gw.data(), chunks, I.begin, I.pos);
}

int map_packed(const HostIndex& I, const mecat_ref_reads* view, const mecat_ref_params* p, long table_budget, mbref::Sink& sink, long* stats, std::string& err,
               std::vector<int32_t>* dump_counts = nullptr, std::vector<int32_t>* dump_rows = nullptr)
{
	const std::vector<uint32_t> rw = words_of(view->vol->pac, view->vol->num_bases);
	HostBackend be;
	be.rfwd = rw.data(); be.roffsz = view->vol->offset_size; be.gfwd = I.gw.data();
	mbref::MapIn in;
	in.R = view->num_reads; in.h_len = view->read_len; in.h_fread = view->fwd_read; in.h_rread = view->rev_read; in.h_rrc = view->rev_is_rc;
	in.seqcount = I.n;
	in.d_fwd = rw.data(); in.d_offsz = view->vol->offset_size; in.d_bad = view->bad; in.nbad = view->num_bad;
	in.d_ibegin = I.begin.data(); in.d_ipos = I.pos.data();
	mbref::Params P;
	P.num_candidates = p->num_candidates; P.num_output = p->num_output; P.want_strings = p->want_strings != 0;
	if (table_budget > 0) P.table_budget = table_budget;
	P.dump_counts = dump_counts; P.dump_rows = dump_rows;
	P.strings_for_printed_only = getenv("MECAT_HARNESS_STRINGS_FOR_PRINTED_ONLY") != NULL;
	P.seed_per_thread = getenv("MECAT_HARNESS_SEED_PER_THREAD") != NULL;      // the scalar bodies instead of the warp-shaped ones
	if (mbref::map_reads(be, in, P, sink)) { err = be.err; return 1; }
	if (stats) { stats[0] += be.tasks_run; stats[1] += be.batches; stats[2] += be.rescue_units; stats[3] += be.string_tasks; }
	return 0;
}

}  // namespace

extern "C" {

void harness_set_tech(int tech) { g_tech = tech; }

// mecat2ref -d reads -r reference -n num_candidates -b num_output -m format through the product's host I/O, stage
// sequence and kernel bodies.  reads_per_call / table_budget force several ABI-sized calls and table batches.
int harness_ref_map(const char* reference_path, const char* reads_path, int num_candidates, int num_output, int format, int reads_per_call,
                    long table_budget, char** text, size_t* bytes, long* stats /* extensions that aligned, seed launches, clipped ends asked about, extensions run with strings */, char* errbuf, int errcap)
{
	const int pack_threads = getenv("MECAT_HARNESS_PACK_THREADS") ? atoi(getenv("MECAT_HARNESS_PACK_THREADS")) : 3;
	auto fail = [&](const std::string& m) { if (errbuf && errcap > 0) snprintf(errbuf, (size_t)errcap, "%s", m.c_str()); return 1; };
	refio::Genome G;
	refio::Reads R;
	std::string err;
	if (!refio::load_genome(reference_path, G, err) || !refio::load_reads(reads_path, R, err)) return fail(err);
	HostIndex I;
	const mecat_ref_genome gv = G.view();
	index_genome(&gv, I);
	std::string out;
	if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
	const int total = (int)R.size();
	if (reads_per_call < 1) reads_per_call = total ? total : 1;
	mecat_ref_params p;
	p.num_candidates = num_candidates; p.num_output = num_output; p.want_strings = format != 1; p.tech = 0;
	for (int first = 0; first < total; first += reads_per_call) {
		const int count = std::min(reads_per_call, total - first);
		refio::ReadBatch B;
		B.build(R, (size_t)first, (size_t)count, pack_threads);
		const mecat_ref_reads view = B.view();
		mbref::Sink sink;
		if (map_packed(I, &view, &p, table_budget, sink, stats, err)) return fail(err);
		refio::format_results(out, G, R.name, first, sink.recs.data(), sink.recs.size(), sink.q.data(), sink.s.data(), format);
	}
	char* o = (char*)malloc(out.size() + 1);
	if (!o) return fail("out of memory");
	memcpy(o, out.data(), out.size());
	o[out.size()] = 0;
	*text = o; *bytes = out.size();
	return 0;
}

// the host twins of mecat_b200_ref_index_build / mecat_b200_ref_map / mecat_b200_ref_index_release: same structures in,
// same records out (lets the CPU suite check the Python packing and formatting of mecat_b200/api.py, and -- through
// tests/ref_abi_shim.cpp -- the command-line driver)
void* harness_ref_index_build(const mecat_ref_genome* g)
{
	HostIndex* I = new HostIndex;
	index_genome(g, *I);
	return I;
}

void harness_ref_index_release(void* idx) { delete (HostIndex*)idx; }

int harness_ref_map_indexed(void* idx, const mecat_ref_reads* reads, const mecat_ref_params* p, mecat_ref_result** results, size_t* n,
                            char** qstrings, char** sstrings, size_t* string_bytes, char* errbuf, int errcap)
{
	mbref::Sink sink;
	std::string err;
	if (map_packed(*(const HostIndex*)idx, reads, p, 0, sink, NULL, err)) { if (errbuf && errcap > 0) snprintf(errbuf, (size_t)errcap, "%s", err.c_str()); return 1; }
	mecat_ref_result* res = (mecat_ref_result*)malloc(sizeof(mecat_ref_result) * (sink.recs.size() ? sink.recs.size() : 1));
	if (!res) return 1;
	if (!sink.recs.empty()) memcpy(res, sink.recs.data(), sizeof(mecat_ref_result) * sink.recs.size());
	*results = res; *n = sink.recs.size(); *string_bytes = sink.q.size();
	*qstrings = sink.q.release(); *sstrings = sink.s.release();
	return 0;
}

// twins of the test hooks mecat_b200_ref_index_export / mecat_b200_ref_raw_candidates
int64_t harness_ref_index_export(void* idx, uint32_t* begin, int32_t* positions)
{
	const HostIndex* I = (const HostIndex*)idx;
	if (begin) memcpy(begin, I->begin.data(), sizeof(uint32_t) * I->begin.size());
	if (positions && !I->pos.empty()) memcpy(positions, I->pos.data(), sizeof(int32_t) * I->pos.size());
	return (int64_t)I->pos.size();
}

int harness_ref_raw_candidates(void* idx, const mecat_ref_reads* reads, const mecat_ref_params* p, int32_t** rows, int32_t** counts, size_t* n)
{
	mbref::Sink sink;
	std::string err;
	std::vector<int32_t> cnt, row;
	if (map_packed(*(const HostIndex*)idx, reads, p, 0, sink, NULL, err, &cnt, &row)) return 1;
	int32_t* r = (int32_t*)malloc(sizeof(int32_t) * (row.size() ? row.size() : 1));
	int32_t* k = (int32_t*)malloc(sizeof(int32_t) * (cnt.size() ? cnt.size() : 1));
	if (!r || !k) return 1;
	if (!row.empty()) memcpy(r, row.data(), sizeof(int32_t) * row.size());
	if (!cnt.empty()) memcpy(k, cnt.data(), sizeof(int32_t) * cnt.size());
	*rows = r; *counts = k; *n = row.size() / 4;
	return 0;
}

int harness_ref_map_packed(const mecat_ref_genome* g, const mecat_ref_reads* reads, const mecat_ref_params* p, mecat_ref_result** results, size_t* n,
                           char** qstrings, char** sstrings, size_t* string_bytes)
{
	void* idx = harness_ref_index_build(g);
	const int rc = harness_ref_map_indexed(idx, reads, p, results, n, qstrings, sstrings, string_bytes, NULL, 0);
	harness_ref_index_release(idx);
	return rc;
}

// the float forms of the DDF test as the reference writes them, next to the integer form of the kernels
int harness_ddf_forms(int a, int b, int bc)
{
	const float len = (float)bc;
	int r = mbref::ddf_close(a, b, bc) ? 1 : 0;
	if (b != 0) {
		const bool f32 = fabs(a / (b * len) - 1) < 0.25;                 // find_location, mecat2ref_aux.cpp:23
		const bool f32d = fabs(a / (b * len) - 1.0) < 0.25;              // insert_loc, mecat2ref_impl_large.cpp:107
		const bool f64 = fabs(a / (b * bc * 1.0) - 1.0) < 0.25;          // the neighbour votes, :520,:551
		r |= (f32 ? 2 : 0) | (f32d ? 4 : 0) | (f64 ? 8 : 0);
	}
	return r;
}

// number of (a, b) pairs in [-amax, amax] x [-bmax, bmax] (b != 0) and strides on which any float form disagrees with the integer form
long harness_ddf_sweep(int amax, int bmax, int bc_lo, int bc_hi)
{
	long bad = 0;
	for (int bc = bc_lo; bc <= bc_hi; ++bc)
		for (int b = -bmax; b <= bmax; ++b) {
			if (!b) continue;
			for (int a = -amax; a <= amax; ++a) {
				const int r = harness_ddf_forms(a, b, bc);
				if (r != 0 && r != 15) ++bad;
			}
		}
	return bad;
}

void harness_free(void* p) { free(p); }

}  // extern "C"
