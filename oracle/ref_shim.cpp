// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" access to the UNMODIFIED reference implementation for function-level parity
// checks.  Nothing is copied: the reference translation unit src/mecat2pw/pw_impl.cpp is
// #included at compile time (so its file-static tunables MAXC / min_kmer_match / ... are
// reachable) and the rest of the reference is linked from objects compiled straight out
// of /root/reference (see oracle/Makefile, target _ref/libmecatref.so).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.

#include "mecat2pw/pw_impl.cpp"          // -I$(REF)/src ; brings in seeding(), get_candidates(), ...
#include "mecat2cns/dw.h"
#include "mecat2cns/reads_correction_aux.h"
#include "mecat2cns/MECAT_AlnGraphBoost.H"

#include <cstdint>
#include <cstring>
#include <vector>

void add_one_seq(volume_t* volume, const char* s, const int size);  // split_database.cpp:104
// diff_gapalign.cpp:107 (defined there, not declared in its header)
int Align(const char* query, const int q_len, const char* target, const int t_len,
          const int band_tolerance, const int get_aln_str, Alignment* align,
          int* V, int* U, DPathData2* d_path, PathPoint* aln_path, const int right_extend);

// defined in mecat_correction.cpp without a header declaration
namespace ns_meap_cns {
void get_effective_ranges(std::vector<MappingRange>& mranges, std::vector<MappingRange>& eranges, const int read_size, const int min_size);
void meap_add_one_aln(const std::string& qaln, const std::string& saln, index_t start_soff, CnsTableItem* cns_table, const char* org_seq);
}

extern "C" {

// ---------------------------------------------------------------- volumes
void* ref_volume_new(int num_bases_cap)
{
	volume_t* v = new_volume_t(0, num_bases_cap);
	v->start_read_id = 0;
	return v;
}

// ASCII read -> volume, exactly like split_raw_dataset (split_database.cpp:250-251).
void ref_volume_add(void* vp, const char* ascii, int size)
{
	volume_t* v = (volume_t*)vp;
	add_one_seq(v, ascii, size);
	++v->curr;
}

void ref_volume_set_start_id(void* vp, int start_read_id) { ((volume_t*)vp)->start_read_id = start_read_id; }
void* ref_volume_load(const char* path) { return load_volume(path); }
void ref_volume_dump(void* vp, const char* path) { dump_volume(path, (volume_t*)vp); }
void ref_volume_free(void* vp) { delete_volume_t((volume_t*)vp); }
int ref_volume_num_reads(void* vp) { return ((volume_t*)vp)->num_reads; }
int ref_volume_num_bases(void* vp) { return ((volume_t*)vp)->curr; }
int ref_volume_start_id(void* vp) { return ((volume_t*)vp)->start_read_id; }
const uint8_t* ref_volume_pac(void* vp) { return ((volume_t*)vp)->data; }
const int* ref_volume_offsets(void* vp) { return (const int*)((volume_t*)vp)->offset_list->offset_list; }
int ref_read_id_from_offset(void* vp, int offset)
{
	return get_read_id_from_offset_list(((volume_t*)vp)->offset_list, offset);
}

// ---------------------------------------------------------------- index (A1)
void* ref_index_create(void* vp, int threads) { return create_ref_index((volume_t*)vp, 13, threads); }
void ref_index_free(void* ip) { destroy_ref_index((ref_index*)ip); }
int ref_index_count(void* ip, uint32_t code) { return ((ref_index*)ip)->kmer_counts[code]; }
const int* ref_index_list(void* ip, uint32_t code) { return ((ref_index*)ip)->kmer_starts[code]; }

// ---------------------------------------------------------------- pw options
void ref_pw_set_options(int num_candidates, int min_align, int min_kmer_match_, int tech)
{
	MAXC = num_candidates;
	min_align_size = min_align;
	min_kmer_match = min_kmer_match_;
	if (tech == TECH_PACBIO) { ddfs_cutoff = ddfs_cutoff_pacbio; min_kmer_dist = 1800; }
	else { ddfs_cutoff = ddfs_cutoff_nanopore; min_kmer_dist = 400; }
}

// ---------------------------------------------------------------- A4 / A5
void ref_insert_loc(short* bucket /* Back_List */, int loc, int seedn)
{
	insert_loc((Back_List*)bucket, loc, seedn, BC);
}
int ref_sizeof_back_list() { return (int)sizeof(Back_List); }

int ref_find_location(int* t_loc, int* t_seedn, int* t_score, int* loc, int k, int* rep_loc, int read_len)
{
	return find_location(t_loc, t_seedn, t_score, loc, k, rep_loc, BC, read_len);
}

// ---------------------------------------------------------------- A2/A3/A6 per read
struct RefPwCtx
{
	volume_t* ref;
	ref_index* ridx;
	SeedingBK* sbk;
	char* read1;
	char* read2;
};

void* ref_pw_ctx_new(void* refvol, void* ridx)
{
	RefPwCtx* c = new RefPwCtx;
	c->ref = (volume_t*)refvol;
	c->ridx = (ref_index*)ridx;
	c->sbk = new SeedingBK(c->ref->curr);
	c->read1 = (char*)malloc(MAX_SEQ_SIZE);
	c->read2 = (char*)malloc(MAX_SEQ_SIZE);
	return c;
}
void ref_pw_ctx_free(void* cp)
{
	RefPwCtx* c = (RefPwCtx*)cp;
	delete c->sbk; free(c->read1); free(c->read2); delete c;
}

// Runs seeding() on one strand and returns the state get_candidates() would see:
// first-touch list, snapshot scores, and for each touched bucket its Back_List.
// Buckets are then reset like get_candidates' tail (pw_impl.cpp:459-463) so the ctx can be reused.
int ref_pw_seeding_dump(void* cp, void* readsvol, int rid, int strand,
                        int* index_list, short* index_score, short* buckets /* used_segs x sizeof(Back_List)/2 */,
                        int cap)
{
	RefPwCtx* c = (RefPwCtx*)cp;
	volume_t* reads = (volume_t*)readsvol;
	int rsize = reads->offset_list->offset_list[rid].size;
	extract_one_seq(reads, rid, c->read1);
	const char* read = c->read1;
	if (strand) { reverse_complement(c->read2, c->read1, rsize); read = c->read2; }
	int used = seeding(read, rsize, c->ridx, c->sbk);
	for (int i = 0; i < used; ++i) {
		int seg = c->sbk->index_list[i];
		if (i < cap) {
			index_list[i] = seg;
			index_score[i] = c->sbk->index_score[i];
			memcpy(buckets + (size_t)i * (sizeof(Back_List) / 2), c->sbk->database + seg, sizeof(Back_List));
		}
		c->sbk->database[seg].score = 0;
		c->sbk->database[seg].index = -1;
	}
	return used;
}

// Full per-read candidate detection: F then R strand into one list (pw_impl.cpp:751-765).
// out = n x 12 ints: loc1 loc2 left1 left2 right1 right2 score num1 num2 readno readstart chain(0=F,1=R)
int ref_pw_candidates(void* cp, void* readsvol, int rid, int chain_as_char, int* out)
{
	RefPwCtx* c = (RefPwCtx*)cp;
	volume_t* reads = (volume_t*)readsvol;
	std::vector<candidate_save> cand(MAXC + 1);
	int rsize = reads->offset_list->offset_list[rid].size;
	extract_one_seq(reads, rid, c->read1);
	reverse_complement(c->read2, c->read1, rsize);
	int n = 0;
	for (int s = 0; s < 2; ++s) {
		const char* read = s ? c->read2 : c->read1;
		char chain = chain_as_char ? (s ? 'R' : 'F') : (char)(s ? REV : FWD);
		int num_segs = seeding(read, rsize, c->ridx, c->sbk);
		n = get_candidates(c->ref, c->sbk, num_segs, rid + reads->start_read_id, rsize, chain, cand.data(), n);
	}
	for (int i = 0; i < n; ++i) {
		int* o = out + 12 * i;
		o[0] = cand[i].loc1; o[1] = cand[i].loc2; o[2] = cand[i].left1; o[3] = cand[i].left2;
		o[4] = cand[i].right1; o[5] = cand[i].right2; o[6] = cand[i].score; o[7] = cand[i].num1;
		o[8] = cand[i].num2; o[9] = cand[i].readno; o[10] = cand[i].readstart;
		o[11] = (cand[i].chain == 'R' || cand[i].chain == REV) ? 1 : 0;
	}
	return n;
}

// ---------------------------------------------------------------- A8-A11: DiffAligner::go
void* ref_diff_new() { return new DiffAligner(0); }
void ref_diff_free(void* a) { delete (DiffAligner*)a; }

// q/t: base codes 0..3, one byte per base.  out[0..5] = ok, qs, qe, ts, te, aln_size.
// Returns ok.  ident through *ident.  Strings (ASCII ACGT-) valid until the next call.
int ref_diff_go(void* ap, const char* q, int qstart, int qsize, const char* t, int tstart, int tsize,
                int min_aln, int* out, double* ident, const char** qstr, const char** tstr)
{
	DiffAligner* a = (DiffAligner*)ap;
	bool ok = a->go(q, qstart, qsize, t, tstart, tsize, min_aln);
	out[0] = ok; out[1] = a->query_start(); out[2] = a->query_end();
	out[3] = a->target_start(); out[4] = a->target_end(); out[5] = a->result->out_store_size;
	*ident = a->calc_ident();
	if (qstr) *qstr = a->query_mapped_string();
	if (tstr) *tstr = a->target_mapped_string();
	return ok;
}

// ---------------------------------------------------------------- nanopore extension (XdropAligner)
void* ref_xdrop_new() { return new XdropAligner(0); }
void ref_xdrop_free(void* a) { delete (XdropAligner*)a; }
int ref_xdrop_go(void* ap, const char* q, int qstart, int qsize, const char* t, int tstart, int tsize,
                 int min_aln, int* out, double* ident, const char** qstr, const char** tstr)
{
	XdropAligner* a = (XdropAligner*)ap;
	bool ok = a->go(q, qstart, qsize, t, tstart, tsize, min_aln);
	out[0] = ok; out[1] = a->query_start(); out[2] = a->query_end();
	out[3] = a->target_start(); out[4] = a->target_end(); out[5] = a->aln_size;
	*ident = a->calc_ident();
	if (qstr) *qstr = a->query_mapped_string();
	if (tstr) *tstr = a->target_mapped_string();
	return ok;
}

// One block: Align() + trim (diff_gapalign.cpp:107, gapalign.cpp:48), for fine-grained checks.
// out = aln_q_e, aln_t_e, dist, aln_str_size, trim_ok, qcnt, tcnt, acnt
void ref_diff_align_block(void* ap, const char* q, int qlen, const char* t, int tlen, int right_extend, int* out)
{
	DiffAligner* a = (DiffAligner*)ap;
	std::fill(a->dynq, a->dynq + a->param.row_size, 0);
	std::fill(a->dynt, a->dynt + a->param.column_size, 0);
	::Align(q, qlen, t, tlen, 0.3 * std::max(qlen, tlen), 400, a->align, a->dynq, a->dynt, a->d_path, a->aln_path, right_extend);
	int qcnt = 0, tcnt = 0, acnt = 0;
	bool trim = trim_mismatch_end(a->align->q_aln_str, a->align->t_aln_str, a->align->aln_str_size, 4, qcnt, tcnt, acnt);
	out[0] = a->align->aln_q_e; out[1] = a->align->aln_t_e; out[2] = a->align->dist;
	out[3] = a->align->aln_str_size; out[4] = trim; out[5] = qcnt; out[6] = tcnt; out[7] = acnt;
}

// ---------------------------------------------------------------- C1/C2: cns GetAlignment
void* ref_cns_drd_new() { return new ns_banded_sw::DiffRunningData(ns_banded_sw::get_sw_parameters_small()); }
void ref_cns_drd_free(void* d) { delete (ns_banded_sw::DiffRunningData*)d; }

// out = ok, qoff, qend, soff, send ; strings are ASCII with '-' gaps, NUL terminated, in caller buffers.
int ref_cns_get_alignment(void* dp, const char* q, int qstart, int qsize, const char* t, int tstart, int tsize,
                          double err, int min_aln, int* out, char* qaln, char* saln, int cap)
{
	static M5Record* m5 = NULL;
	if (!m5) m5 = NewM5Record(MAX_SEQ_SIZE);
	bool ok = ns_banded_sw::GetAlignment(q, qstart, qsize, t, tstart, tsize,
	                                     (ns_banded_sw::DiffRunningData*)dp, *m5, err, min_aln);
	out[0] = ok;
	if (ok) {
		out[1] = m5qoff(*m5); out[2] = m5qend(*m5); out[3] = m5soff(*m5); out[4] = m5send(*m5);
		strncpy(qaln, m5qaln(*m5), cap - 1); qaln[cap - 1] = 0;
		strncpy(saln, m5saln(*m5), cap - 1); saln[cap - 1] = 0;
	}
	return ok;
}

// C4: normalize_gaps (reads_correction_aux.cpp:3)
int ref_normalize_gaps(const char* qstr, const char* tstr, int n, int push, char* qout, char* tout, int cap)
{
	std::string qn, tn;
	normalize_gaps(qstr, tstr, n, qn, tn, push != 0);
	if ((int)qn.size() + 1 > cap) return -1;
	memcpy(qout, qn.c_str(), qn.size() + 1);
	memcpy(tout, tn.c_str(), tn.size() + 1);
	return (int)qn.size();
}

// C6: get_effective_ranges (mecat_correction.cpp:118-153; defined there without a header declaration)
int ref_effective_ranges(const int* in, int n, int read_size, int min_size, int* out, int cap)
{
	std::vector<MappingRange> m, e;
	for (int i = 0; i < n; ++i) m.push_back(MappingRange(in[2 * i], in[2 * i + 1]));
	ns_meap_cns::get_effective_ranges(m, e, read_size, min_size);
	if ((int)e.size() > cap) return -1;
	for (size_t i = 0; i < e.size(); ++i) { out[2 * i] = e[i].start; out[2 * i + 1] = e[i].end; }
	return (int)e.size();
}

// C4 + C5: normalize_gaps then meap_add_one_aln into a table of `positions` items.  out = positions x {base, mat, ins, del};
// the normalised strings go to qout / tout.  Returns the normalised length.
int ref_normalize_and_vote(const char* qstr, const char* tstr, int n, int soff, int positions, unsigned char* out, char* qout, char* tout, int cap)
{
	std::string qn, tn;
	normalize_gaps(qstr, tstr, n, qn, tn, true);
	std::vector<CnsTableItem> table((size_t)positions);
	ns_meap_cns::meap_add_one_aln(qn, tn, soff, table.data(), NULL);
	for (int i = 0; i < positions; ++i) { out[4 * i] = (unsigned char)table[i].base; out[4 * i + 1] = table[i].mat_cnt; out[4 * i + 2] = table[i].ins_cnt; out[4 * i + 3] = table[i].del_cnt; }
	if ((int)qn.size() + 1 > cap) return -1;
	memcpy(qout, qn.c_str(), qn.size() + 1);
	memcpy(tout, tn.c_str(), tn.size() + 1);
	return (int)qn.size();
}

// C7: one AlnGraphBoost region graph exactly as meap_cns_one_indel drives it (mecat_correction.cpp:63-78)
int ref_poa_consensus(int blen, int naln, const char* const* q, const char* const* t, const int* start, int min_weight, char* out, int cap)
{
	ns_meap_cns::AlnGraphBoost ag(blen);
	for (int i = 0; i < naln; ++i) ag.addAln(std::string(q[i]), std::string(t[i]), (size_t)start[i]);
	ag.mergeNodes();
	std::string cns;
	ag.consensus(min_weight, cns);
	if ((int)cns.size() > cap) return -1;
	memcpy(out, cns.data(), cns.size());
	return (int)cns.size();
}

} // extern "C"
