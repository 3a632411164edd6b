/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the MECAT hot path, written from the
 * reference's behaviour, not from its text.  Every function cites the reference
 * file:line it follows.  It is pinned against the real reference in two ways:
 *   - tests/test_oracle_vs_ref.py compares it function by function with
 *     oracle/_ref/libmecatref.so (the unmodified reference, built by oracle/Makefile);
 *   - tests/golden/ holds outputs of the unmodified reference binaries on seeded inputs
 *     (tests/golden/make_golden.py) which the oracle must reproduce bit for bit.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 * The product library (mecat_b200/csrc) never links or calls anything in oracle/.
 */
#ifndef MECAT_ORACLE_H
#define MECAT_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same plain-data view of a packed volume as include/mecat_b200.h (volume_t on disk,
 * split_database.cpp:136-153): offset_size = {offset,size} pairs, pac = 2 bit/base. */
typedef struct {
	int32_t num_reads, num_bases, start_read_id;
	const int32_t* offset_size;
	const uint8_t* pac;
} orc_volume;

typedef struct {
	int32_t task, num_candidates, min_align_size, min_kmer_match, tech;
} orc_pw_params;

/* ---- FASTA -> volumes (split_database.cpp:222-266) : returns malloc'ed arrays */
int orc_pack_reads(const char* const* seqs, const int32_t* lens, int n, int32_t** offset_size, uint8_t** pac,
                   int32_t* num_bases);

/* ---- A1 index */
void* orc_index_build(const orc_volume* v);
void orc_index_free(void* idx);
int orc_index_lookup(const void* idx, uint32_t code, const int32_t** list);
int64_t orc_index_num_kmers(const void* idx);

/* ---- A2-A4 seeding of one strand.  Output per touched bucket, first-touch order:
 * seg[i], index_score[i], and a row of 82 shorts: score, loczhi[40], seedno[40], pad. */
int orc_seeding(const void* idx, const orc_volume* ref, const orc_volume* reads, int rid, int strand,
                int32_t* seg, int16_t* index_score, int16_t* rows, int cap);

/* ---- A4/A5 building blocks */
void orc_insert_loc(int16_t* score, int16_t* loczhi, int16_t* seedno, int loc, int seedn);
int orc_find_location(const int* t_loc, const int* t_seedn, int* t_score, int* loc, int k, int* rep_loc,
                      int read_len);

/* ---- A2-A6 candidates of one read: n x 12 ints
 * loc1 loc2 left1 left2 right1 right2 score num1 num2 readno readstart chain */
int orc_pw_candidates(const void* idx, const orc_volume* ref, const orc_volume* reads, int rid,
                      const orc_pw_params* p, int32_t* out);

/* ---- A8-A11 extension (DiffAligner::go).  q,t: codes 0..3.  out = ok qs qe ts te aln_size matches.
 * qstr/tstr (optional, cap bytes each) receive the ASCII alignment. */
int orc_diff_go(const char* q, int qstart, int qsize, const char* t, int tstart, int tsize, int min_aln,
                int32_t* out, double* ident, char* qstr, char* tstr, int cap);
/* ---- nanopore extension (XdropAligner::go, common/xdrop_gapalign.cpp:10-439); same conventions as orc_diff_go,
 * ok = qend - qoff >= min_aln, out[5] = aligned columns, out[6] = columns with equal letters */
int orc_xdrop_go(const char* q, int qstart, int qsize, const char* t, int tstart, int tsize, int min_aln,
                 int32_t* out, double* ident, char* qstr, char* tstr, int cap);
/* one block: aln_q_e aln_t_e dist aln_str_size trim_ok qcnt tcnt acnt */
void orc_diff_align_block(const char* q, int qlen, const char* t, int tlen, int right_extend, int32_t* out);

/* ---- A7/A12 whole tile (process_one_volume for one (index volume, query volume) pair).
 * task 0: records = ExtensionCandidate (13 int32); task 1: M4Record (104 B).
 * Records come out read by read, inside a read in the reference's own order.
 * Result buffer is malloc'ed; free with orc_free. */
int orc_pw_tile(const orc_volume* ref, const orc_volume* reads, const orc_pw_params* p, int threads,
                void** records, size_t* n);
void orc_free(void* p);

/* ---- C1/C2 consensus-flavour extension (mecat2cns/dw.cpp GetAlignment).
 * out = ok qoff qend soff send ; strings ASCII with '-' */
int orc_cns_get_alignment(const char* q, int qstart, int qsize, const char* t, int tstart, int tsize,
                          double err, int min_aln, int32_t* out, char* qaln, char* saln, int cap);
/* ---- C4 */
int orc_normalize_gaps(const char* qstr, const char* tstr, int n, int push, char* qout, char* tout, int cap);

/* ---- mecat2ref end to end (next row of the scope table; no CUDA path yet): the text `mecat2ref -m format` writes,
 * format 0 = ref (header + two alignment strings), 1 = m4.  malloc'ed, free with orc_free. */
int orc_ref_map(const char* reference_path, const char* reads_path, int num_candidates, int num_output, int format,
                char** text, size_t* bytes);

int orc_ref_map_x(const char* reference_path, const char* reads_path, int num_candidates, int num_output, int format, int tech,
                  char** text, size_t* bytes);

#ifdef __cplusplus
}
#endif
#endif
