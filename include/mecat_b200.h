/* include/mecat_b200.h -- C ABI of the B200-native MECAT overlap hot path.
 *
 * The reference (xiaochuanle/MECAT) has no plugin/FFI layer; its seams are ordinary C++
 * functions called once per read.  A per-read call is too fine for a GPU, so the boundary
 * sits at the batch seams its drivers already have.  Each entry
 * point names the reference interface it replaces.  Plain pointers and sizes only.
 *
 * Conventions: one context per device, used from one host thread; 0 = success, any other
 * value = failure with text in mecat_b200_last_error(); inputs are borrowed for the call;
 * outputs are library-owned host buffers released with mecat_b200_free().  The library
 * never aborts and has no CPU fallback: without a CUDA device mecat_b200_init fails.
 */
#ifndef MECAT_B200_H
#define MECAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MECAT_B200_ABI_VERSION 1

typedef struct mecat_b200_ctx mecat_b200_ctx;

/* A 2-bit packed read volume exactly as the reference keeps it in `wrk/volN`
 * (volume_t, src/common/split_database.h:18-24; file layout split_database.cpp:136-153):
 * offset_size = num_reads x {offset,size}; pac holds (num_bases+3)/4 bytes, base i in byte
 * i>>2 at shift ((~i)&3)<<1 (packed_db.h:98-107); one zero pad base follows every read. */
typedef struct {
	int32_t num_reads, num_bases, start_read_id;
	const int32_t* offset_size;
	const uint8_t* pac;
} mecat_volume;

/* Mirrors options_t (src/mecat2pw/pw_options.h:9-21): -j, -n, -a, -k, -x. */
typedef struct {
	int32_t task;            /* 0 = candidates (.can), 1 = overlaps (.m4)            */
	int32_t num_candidates;  /* -n, default 100                                      */
	int32_t min_align_size;  /* -a, default 2000 (pacbio)                            */
	int32_t min_kmer_match;  /* -k, default 4 (pacbio)                               */
	int32_t tech;            /* -x, 0 = pacbio, 1 = nanopore (XdropAligner, min_kmer_dist 400; pw_impl.cpp:638-642,843-849) */
} mecat_pw_params;

/* ExtensionCandidate, src/common/alignment.h:8-13 (13 x int32 = 52 bytes). */
typedef struct {
	int32_t qdir, qid, qext, qsize, qoff, qend;
	int32_t sdir, sid, sext, ssize, soff, send;
	int32_t score;
} mecat_candidate;

/* M4Record, src/common/alignment.h:21-37 (104 bytes; idx_t = int64). */
typedef struct {
	int64_t qid, sid;
	double ident;
	int32_t vscore, qdir;
	int64_t qoff, qend, qsize;
	int32_t sdir, pad_;
	int64_t soff, send, ssize, qext, sext;
} mecat_m4;

/* One gapped-extension request: GapAligner::go(query, qstart, qsize, target, tstart,
 * tsize, min_aln) (src/common/gapalign.h:14-16) with both sequences named by read index
 * inside an uploaded volume.  qstrand 1 = the query is the reverse complement of the
 * read and qstart is in that orientation (pw_impl.cpp:676-688). */
typedef struct {
	int32_t qread, qstrand, qstart;
	int32_t sread, sstart;
} mecat_extend_task;

/* Results of DiffAligner::go + accessors (src/common/diff_gapalign.cpp:295-349,
 * diff_gapalign.h:159-181): ok = aligned columns >= min_aln; ident = 100*matches/columns. */
typedef struct {
	int32_t ok, qstart, qend, sstart, send, columns, matches, pad_;
	double ident;
} mecat_extend_result;

/* Gapped extension WITH alignment strings.  Like mecat_extend_task plus an optional window on the
 * subject (swin_len > 0: the target is bases [swin_off, swin_off+swin_len) of read sread and
 * sstart is relative to the window -- mecat2ref cuts such a window out of the reference,
 * src/mecat2ref/mecat2ref_aux.cpp:86-121). */
typedef struct {
	int32_t qread, qstrand, qstart;
	int32_t sread, sstart;
	int32_t swin_off, swin_len;
} mecat_align_task;

/* policy 0: DiffAligner::go accessors; policy 1: the M5Record fields GetAlignment fills
 * (src/mecat2cns/dw.cpp:482-553: coordinates and strings trimmed to a 4-match run at both ends).
 * str_offset: offset of this task's NUL-terminated strings in both string arenas, -1 if !ok. */
typedef struct {
	int32_t ok, qstart, qend, sstart, send, columns, matches, pad_;
	double ident;
	int64_t str_offset;
} mecat_align_result;

/* Device time per kernel (milliseconds, CUDA events on the context's own stream, summed over
 * launches since the last reset) and traffic counters.  Indices into kernel_ms / kernel_launches: */
enum {
	MECAT_K_ORIENT = 0,  /* volume re-layout (fwd / reversed 2-bit words)            */
	MECAT_K_COUNT = 1,   /* index: k-mer histogram                                   */
	MECAT_K_SCAN = 2,    /* index: cutoff + exclusive scan                           */
	MECAT_K_FILL = 3,    /* index: position scatter                                  */
	MECAT_K_SORT = 4,    /* index: per-list ordering                                 */
	MECAT_K_SEED = 5,    /* seeding: hit streaming, bucket records                   */
	MECAT_K_WALK = 6,    /* DDF scoring + candidate walk                             */
	MECAT_K_MERGE = 7,   /* per-read candidate merge + record assembly               */
	MECAT_K_EXTEND = 8,  /* O(nd) diff extension                                     */
	MECAT_K_FINAL = 9,   /* extension result assembly                                */
	MECAT_K_CNS_ACCEPT = 10,   /* cns: accept loop of every read (C3)                */
	MECAT_K_CNS_NORMVOTE = 11, /* cns: normalize_gaps + pile-up votes (C4-C5)        */
	MECAT_K_CNS_SEGMENT = 12,  /* cns: effective ranges, covered runs (C6)           */
	MECAT_K_CNS_REGION = 13,   /* cns: position flags, anchors, ambiguous regions    */
	MECAT_K_CNS_POA = 14,      /* cns: one partial-order graph per region (C7)       */
	MECAT_K_CNS_ASSEMBLE = 15, /* cns: corrected bases of every segment              */
	MECAT_K_REF_COUNT = 16,    /* ref: index hits per strand (sizes the block tables) */
	MECAT_K_REF_SEED = 17,     /* ref: seeding, DDF scoring, candidate walk           */
	MECAT_K_REF_RESCUE = 18,   /* ref: candidates beyond clipped alignment ends       */
	MECAT_K_ASM_INDEX = 19,    /* asmpw: k-mer index of the subject file's text       */
	MECAT_K_ASM_SEED = 20,     /* asmpw: hit counts, block tables, candidate walk, merge */
	MECAT_K_ASM_EXTEND = 21,   /* asmpw: chunked O(nd) alignment, gap shifting, records */
	MECAT_K_NUM = 22
};
typedef struct {
	float kernel_ms[MECAT_K_NUM];
	int64_t kernel_launches[MECAT_K_NUM];
	float h2d_ms, d2h_ms, host_ms, total_ms;
	float wall_index_ms, wall_seed_ms, wall_extend_ms, wall_other_ms;   /* host wall clock spent inside each phase */
	int64_t h2d_bytes, d2h_bytes;
	int64_t num_hits, num_candidates, num_extend_blocks, index_kmers, index_bases, num_records;
	int64_t num_extend_cells;    /* furthest-point cell updates of the O(nd) extension (k_extend_lanes) */
	int64_t num_extend_spills;   /* extension chains finished by the wide-band kernel                  */
} mecat_b200_stats;

/* ---- lifetime ---------------------------------------------------------------------- */
int mecat_b200_abi_version(void);
int mecat_b200_device_count(void);
/* nccl_comm_or_null: reserved for the multi-GPU block rotation; unused at N=1. */
int mecat_b200_init(mecat_b200_ctx** ctx, int device, void* nccl_comm_or_null);
void mecat_b200_destroy(mecat_b200_ctx* ctx);
const char* mecat_b200_last_error(mecat_b200_ctx* ctx);
void mecat_b200_free(mecat_b200_ctx* ctx, void* p);
int mecat_b200_get_stats(mecat_b200_ctx* ctx, mecat_b200_stats* out);
int mecat_b200_reset_stats(mecat_b200_ctx* ctx);

/* ---- on-disk volumes (host only, byte-compatible with the reference) -----------------------
 * replaces split_raw_dataset / dump_volume (src/common/split_database.cpp:222-266,136-153):
 * FASTA/FASTQ -> wrk_dir/vol0..N-1 + wrk_dir/fileindex.txt.  max_volume_bases <= 0 selects the
 * reference's 2 140 000 000-base cap (MCS, split_database.h:6). */
int mecat_b200_split_dataset(const char* reads_path, const char* wrk_dir, int64_t max_volume_bases,
                             int* num_volumes, char* err, int err_cap);
/* All reads of a FASTA/FASTQ file as one packed volume in host memory; replaces PackedDB::load_fasta_db
 * (src/common/packed_db.cpp:194) for mecat2cns.  Same bytes as vol0 of mecat_b200_split_dataset; fails when the reads
 * need more than one volume.  Release with mecat_b200_volume_unload. */
int mecat_b200_volume_from_fasta(const char* reads_path, mecat_volume* out, char* err, int err_cap);

/* replaces load_volume / delete_volume_t (split_database.cpp:156-181,95-101). */
/* The whole read set as volumes in host memory, cut like mecat_b200_split_dataset cuts its files (max_volume_bases <= 0: the
 * reference's 2.14 Gbase); for read sets beyond one volume (mecat_b200_cns_reads_multi).  *vols: malloc'ed array. */
int mecat_b200_volumes_from_fasta(const char* reads_path, int64_t max_volume_bases, mecat_volume** vols, int* num_volumes,
                                  char* err, int err_cap);
void mecat_b200_volumes_unload(mecat_volume* vols, int num_volumes);
int mecat_b200_volume_load(const char* path, mecat_volume* out);
void mecat_b200_volume_unload(mecat_volume* v);

/* ---- volumes resident in HBM --------------------------------------------------------
 * replaces load_volume (split_database.cpp:156-181) + extract_one_seq/reverse_complement
 * (split_database.cpp:122-133, pw_impl.cpp:69-81): the packed bases are copied to the
 * device once and re-laid out in both walking directions. */
int mecat_b200_volume_upload(mecat_b200_ctx* ctx, const mecat_volume* v, void** dvol);
int mecat_b200_volume_release(mecat_b200_ctx* ctx, void* dvol);
/* Same, but the packed bytes are already in device memory (a query volume received from a peer
 * GPU over NCCL in the block rotation): only the small offset table comes from the host. */
int mecat_b200_volume_from_device(mecat_b200_ctx* ctx, int32_t num_reads, int32_t num_bases, int32_t start_read_id,
                                  const int32_t* host_offset_size, const void* device_pac, void** dvol);

/* 2-bit packing on the device: replaces add_one_seq / PackedDB::set_char for a whole volume (src/common/split_database.cpp:
 * 103-119, src/common/packed_db.h:98-101; same bytes, including how codes above 3 -- N, other IUPAC letters -- spill inside
 * their byte).  text: letters of the reads (host memory, e.g. the mapped FASTA file); read i has offset_size[2i+1] letters
 * starting at text[src_offset[i]] (contiguous: a read spread over several lines is copied together by the caller) and lands
 * at base offset_size[2i] of the volume -- the offsets the reference's splitting rule gives (one pad base after every read,
 * split_database.cpp:240-259).  The volume stays resident (*dvol); pac_out, when not NULL, receives the (num_bases+3)/4
 * packed bytes for the volume file. */
int mecat_b200_volume_from_text(mecat_b200_ctx* ctx, const char* text, size_t text_bytes, const int64_t* src_offset,
                                const int32_t* offset_size, int32_t num_reads, int32_t num_bases, int32_t start_read_id,
                                uint8_t* pac_out, void** dvol);

/* ---- A1: k-mer index of an index volume ----------------------------------------------
 * replaces create_ref_index (src/common/lookup_table.cpp:64-160). */
int mecat_b200_index_build(mecat_b200_ctx* ctx, void* dvol_ref, void** index);
int mecat_b200_index_release(mecat_b200_ctx* ctx, void* index);
/* The same build in two stages, for several GPUs that each own a slice [code_lo, code_hi) of the
 * 2^26 k-mer codes (multiples of 256): count_part histograms the slice; the caller all-gathers the
 * count slices (device array from index_device_arrays); finish_part scans all codes and fills /
 * orders the positions of its slice, which the caller then broadcasts. */
int mecat_b200_index_count_part(mecat_b200_ctx* ctx, void* dvol_ref, uint32_t code_lo, uint32_t code_hi, void** index);
int mecat_b200_index_finish_part(mecat_b200_ctx* ctx, void* dvol_ref, void* index, uint32_t code_lo, uint32_t code_hi);
/* device addresses of the index arrays: counts (2^26 uint32, only between the two stages, else NULL),
 * begin (2^26+1 uint32), positions (*num_kmers int32) */
int mecat_b200_index_device_arrays(mecat_b200_ctx* ctx, void* index, void** d_counts, void** d_begin, void** d_positions,
                                   int64_t* num_kmers);
/* test hook: number of kept k-mer starts and (optionally) the CSR arrays copied to the host:
 * begin = 2^26+1 uint32, positions = *num_kmers int32; either pointer may be NULL. */
int mecat_b200_index_export(mecat_b200_ctx* ctx, void* index, int64_t* num_kmers, uint32_t* begin,
                            int32_t* positions);

/* ---- A2-A12: one (index volume, query volume) tile -----------------------------------
 * replaces the body of process_one_volume for one query volume (pw_impl.cpp:859-879):
 * candidate_detect (task 0, pw_impl.cpp:720-818) or pairwise_mapping (task 1, :623-718).
 * Device-resident inputs; records come back read by read, inside a read in the
 * reference's own order.  *records = mecat_candidate[] or mecat_m4[]. */
int mecat_b200_pw_tile(mecat_b200_ctx* ctx, void* index, void* dvol_ref, void* dvol_reads,
                       const mecat_pw_params* p, void** records, size_t* n);

/* Same for the query reads [read_begin, read_end) only (read_end < 0 = all): lets several GPUs
 * share one tile (mirror-paired indices). */
int mecat_b200_pw_tile_range(mecat_b200_ctx* ctx, void* index, void* dvol_ref, void* dvol_reads,
                             const mecat_pw_params* p, int read_begin, int read_end, void** records, size_t* n);

/* Same tile, but the result comes back as the LINES of the reference's output file -- operator<<(ExtensionCandidate)
 * for task 0, output_m4record for task 1 (src/common/alignment.cpp:18-32,58-78, src/mecat2pw/pw_impl.cpp:509-531;
 * gapped = `-g 1`: with the two extension points) -- written on the device, so the records never visit the host.
 * *text: malloc'ed, NUL terminated, *bytes long; release with mecat_b200_free. */
int mecat_b200_pw_tile_text(mecat_b200_ctx* ctx, void* index, void* dvol_ref, void* dvol_reads, const mecat_pw_params* p,
                            int gapped, char** text, size_t* bytes, size_t* num_records);
/* The same text for records the caller holds: kind 0 = mecat_candidate[], kind 1 = mecat_m4[]. */
int mecat_b200_records_text(mecat_b200_ctx* ctx, int kind, int gapped, const void* records, size_t n, char** text, size_t* bytes);

/* Same, with host buffers in and out (upload + index build + tile + download): the
 * end-to-end call a host driver makes per tile when nothing is cached. */
int mecat_b200_pw_candidates(mecat_b200_ctx* ctx, const mecat_volume* ref, const mecat_volume* reads,
                             const mecat_pw_params* p, mecat_candidate** ec, size_t* n);
int mecat_b200_pw_overlaps(mecat_b200_ctx* ctx, const mecat_volume* ref, const mecat_volume* reads,
                           const mecat_pw_params* p, mecat_m4** m4, size_t* n);

/* test hook: raw candidate_save lists of get_candidates (pw_impl.cpp:288-465), 12 ints per
 * candidate (loc1 loc2 left1 left2 right1 right2 score num1 num2 readno readstart chain),
 * counts[r] candidates for read r, rows concatenated in read order. */
int mecat_b200_pw_raw_candidates(mecat_b200_ctx* ctx, void* index, void* dvol_ref, void* dvol_reads,
                                 const mecat_pw_params* p, int32_t** rows, int32_t** counts, size_t* n);

/* ---- A8-A11: batched gapped extension --------------------------------------------------
 * replaces GapAligner::go per candidate (pw_impl.cpp:688, mecat2ref_aux.cpp:152).
 * policy 0 = pw/ref flavour (common/diff_gapalign.cpp); policy 2 = the nanopore flavour, XdropAligner::go
 * (common/xdrop_gapalign.cpp:10-439: ok = query span >= min_align_size, matches = columns with equal letters). */
int mecat_b200_extend_batch(mecat_b200_ctx* ctx, int policy, void* dvol_query, void* dvol_subject,
                            const mecat_extend_task* tasks, size_t ntasks, int min_align_size,
                            mecat_extend_result** results);

/* ---- A8-A11 / R1 / C1-C2: batched gapped extension with alignment strings ------------------
 * policy 0 replaces GapAligner::go + query/target_mapped_string per candidate
 *   (src/mecat2ref/mecat2ref_aux.cpp:152-163, src/common/diff_gapalign.cpp:295-349);
 * policy 1 replaces ns_banded_sw::GetAlignment per candidate with error rate `err`
 *   (src/mecat2cns/mecat_correction.cpp:424, src/mecat2cns/dw.cpp:482-553).
 * policy 2 replaces XdropAligner::go + mapped strings (src/common/xdrop_gapalign.cpp:351-439), the aligner of `-x 1`;
 * Strings are ASCII over ACGT- ; *qstrings / *sstrings hold *string_bytes bytes each. */
int mecat_b200_align_batch(mecat_b200_ctx* ctx, int policy, double err, void* dvol_query, void* dvol_subject,
                           const mecat_align_task* tasks, size_t ntasks, int min_align_size,
                           mecat_align_result** results, char** qstrings, char** sstrings, size_t* string_bytes);

/* ---- C1-C7: consensus of a set of reads (mecat2cns -i 0) -----------------------------------
 * replaces consensus_one_partition_can / reads_correction_func_can / consensus_one_read_can_pacbio
 * (src/mecat2cns/reads_correction_can.cpp:21-86, src/mecat2cns/mecat_correction.cpp:389-450).
 * ec: candidates already normalised like the partition files (the read to correct is `sid`,
 * sdir == 0; src/mecat2cns/overlaps_partition.cpp:141-165), any order; reads with fewer than
 * min_cov candidates or shorter than 0.95 * min_size are skipped like the reference does.
 * dvol_reads must hold every read named by ec (read id = index + start_read_id).
 * Mirrors ConsensusOptions (src/mecat2cns/options.h:9-23): -r, -a, -c, -l, -x. */
typedef struct {
	double min_mapping_ratio;   /* -r, default 0.9  */
	int32_t min_align_size;     /* -a, default 2000 */
	int32_t min_cov;            /* -c, default 6    */
	int64_t min_size;           /* -l, default 5000 */
	int32_t tech;               /* -x, 0 = pacbio; 1 = nanopore: consensus_one_read_can_nanopore (mecat_correction.cpp:453-512:
	                               error rate 0.20, up to 100 alignments per read, the whole read as the one effective range);
	                               its defaults are -r 0.4 -a 400 -c 6 -l 2000 (options.cpp:21-29) */
	int32_t input_type;         /* -i.  0 = candidates (`.can` of mecat2pw -j 0): consensus_one_read_can_* -- trial order by score, one
	                               alignment per partner, mapping-ratio and coverage gates.  1 = overlaps (`.m4` of mecat2pw -j 1 -g 1):
	                               consensus_one_read_m4_* (mecat_correction.cpp:242-360) -- `ec` is ONE partition of
	                               partition_m4records in the order it was written (m4_to_candidate of both normalised
	                               directions, record by record); the library orders it like the reference does (std::sort by sid,
	                               the 60 / 100 largest overlaps of a read by std::sort) and uses every alignment that succeeds
	                               (nanopore: that also passes the mapping-ratio test) */
} mecat_cns_params;

/* CnsResult (src/common/alignment.h): corrected piece [beg, end) of read id; sequence at
 * seqs[seq_offset .. seq_offset + seq_len) (ASCII, not NUL separated). */
typedef struct {
	int64_t id, beg, end, seq_offset, seq_len;
} mecat_cns_piece;

int mecat_b200_cns_reads(mecat_b200_ctx* ctx, void* dvol_reads, const mecat_candidate* ec, size_t nec,
                         const mecat_cns_params* p, mecat_cns_piece** pieces, size_t* npieces, char** seqs,
                         size_t* seq_bytes);

/* The same for reads that span several resident volumes (the read ids of volume v+1 continue those of volume v): any read
 * set the splitter can cut, e.g. the 8 volumes of BASELINE configs[4].  The reference keeps the whole data set in one
 * PackedDB (src/common/packed_db.cpp:194); here the reads a run of templates needs are gathered into a working volume on
 * the device.  Results are independent of how the reads are cut into volumes. */
int mecat_b200_cns_reads_multi(mecat_b200_ctx* ctx, void* const* dvols_reads, int num_volumes, const mecat_candidate* ec, size_t nec,
                               const mecat_cns_params* p, mecat_cns_piece** pieces, size_t* npieces, char** seqs,
                               size_t* seq_bytes);

/* Trial order of one read's candidates (CmpExtensionCandidateByScore, src/mecat2cns/mecat_correction.cpp:362-370,409);
 * host only.  mecat_b200_cns_reads applies it itself; exported so that callers feeding mecat_b200_align_batch can
 * reproduce the order. */
void mecat_b200_cns_sort_candidates(mecat_candidate* c, int n);
/* ---- mecat2ref: reads against a reference genome ---------------------
 * The reference genome as creat_ref_index keeps it (src/mecat2ref/mecat2ref_impl_large.cpp:133-271): all sequences
 * concatenated (REFSEQ), here 2-bit packed like mecat_volume.pac with every letter other than ACGT packed as A
 * (extract_sequences aligns them as A, src/mecat2ref/mecat2ref_aux.cpp:86-121), plus the maximal runs of ACGT letters:
 * the index holds the 13-mers inside a run only (:199-206).  Fewer than 2^31 - 2^20 bases. */
typedef struct {
	int64_t num_bases;
	const uint8_t* pac;
	int32_t num_runs;
	const int64_t* run_start_len;   /* num_runs x {start, length}, ascending, disjoint */
} mecat_ref_genome;

/* A batch of reads for mecat2ref.  vol holds the strands 2-bit packed (letters other than ACGT, either case, as A).
 * Read r has read_len[r] bases; its forward strand is volume read fwd_read[r]; its reverse strand is the reverse
 * complement of volume read rev_read[r] when rev_is_rc[r], else volume read rev_read[r] as packed (the reference
 * complements upper-case ACGT only, mecat2ref_impl_large.cpp:374-400, so a read with other letters carries its reverse
 * strand explicitly).  bad: ascending base offsets inside vol of letters that are not upper-case ACGT; a k-mer covering
 * one is not looked up (transnum_buchang, :64-90).  Strands taken by reverse complement must be free of them. */
typedef struct {
	int32_t num_reads;
	const mecat_volume* vol;
	const int32_t* read_len;
	const int32_t* fwd_read;
	const int32_t* rev_read;
	const int32_t* rev_is_rc;
	int64_t num_bad;
	const int64_t* bad;
} mecat_ref_reads;

/* Mirrors meap_ref_options (src/mecat2ref/mecat2ref.cpp:22-50): -n, -b, -x; want_strings = the output format prints
 * the alignment strings (-m 0). */
typedef struct {
	int32_t num_candidates;   /* -n, default 10 */
	int32_t num_output;       /* -b, default 10 */
	int32_t want_strings;
	int32_t tech;             /* -x, 0 = pacbio (DiffAligner), 1 = nanopore (XdropAligner; mecat2ref_impl_large.cpp:329-332) */
} mecat_ref_params;

/* TempResult (src/mecat2ref/mecat2ref_aux.h:16-25) of one printed alignment: read = index inside the batch, dir 0 = F
 * 1 = R, [qb, qe) on the strand that aligned, [sb, se) on the concatenated reference; columns / matches of the
 * alignment strings, which start at str_offset in both string blobs (NUL terminated; -1 without want_strings). */
typedef struct {
	int32_t read, dir, vscore, qb, qe, qs;
	int64_t sb, se;
	int32_t columns, matches;
	int64_t str_offset;
} mecat_ref_result;

/* replaces creat_ref_index (mecat2ref_impl_large.cpp:133-271): upload + k-mer index of the genome */
int mecat_b200_ref_index_build(mecat_b200_ctx* ctx, const mecat_ref_genome* g, void** refidx);
int mecat_b200_ref_index_release(mecat_b200_ctx* ctx, void* refidx);
/* replaces reference_mapping for a batch of reads (mecat2ref_impl_large.cpp:274-891: both seeding passes, candidate
 * selection, extend_candidate, rescue_clipped_align, output_results).  Records of a read are adjacent and in the
 * reference's order, at most num_output per read. */
int mecat_b200_ref_map(mecat_b200_ctx* ctx, void* refidx, const mecat_ref_reads* reads, const mecat_ref_params* p,
                       mecat_ref_result** results, size_t* n, char** qstrings, char** sstrings, size_t* string_bytes);

/* test hooks: the genome's k-mer index like mecat_b200_index_export (0-based k-mer starts; the reference's databaseindex
 * holds them + 1), and the candidate lists the first seeding pass leaves per strand (counts = 2 per read: forward,
 * reverse; rows = 4 ints per candidate: loc1 loc2 score chain, the fields of `candidate_save` extend_candidate reads;
 * mecat2ref_impl_large.cpp:463-614). */
int mecat_b200_ref_index_export(mecat_b200_ctx* ctx, void* refidx, int64_t* num_kmers, uint32_t* begin, int32_t* positions);
int mecat_b200_ref_raw_candidates(mecat_b200_ctx* ctx, void* refidx, const mecat_ref_reads* reads, const mecat_ref_params* p,
                                  int32_t** rows, int32_t** counts, size_t* n);

/* ---- mecat2asmpw / mecat2trimpw: the overlappers mecat2canu runs on corrected reads ------------
 * These programs (mecat2canu/src/mecat2asmpw/mecat2asmpw.c, mecat2trimpw.c and their *50 twins) work on letters: a file
 * of reads is one text with a NUL behind every read (load_read :345-372, load_fastq :1000-1029), read r at read_start[r],
 * numbered first_read_id + r (the `-b` value of the file's line in `ovlprep`).  Letters are upper-cased by the library
 * like the loaders do; everything other than ACGT ends a k-mer and is compared as it is.  A read must be shorter than
 * 100 000 letters (RM, :17: the reference's fixed buffers) and the text shorter than 2^31 - 4 000. */
typedef struct {
	const char* text;
	int64_t num_letters;          /* including the NUL behind every read */
	int32_t num_reads, first_read_id;
	const int32_t* read_start;
	const int32_t* read_len;
} mecat_asm_reads;

typedef struct {
	int32_t variant;              /* 0 = mecat2asmpw, 1 = mecat2trimpw (gate :640, printed score :942-943) */
	int32_t max_candidates;       /* MAXC :23: 100, the *50 programs 50 */
} mecat_asm_params;

/* One printed line (:944-945): "sread qread score 100 0 sbeg send slen strand qbeg qend qlen", score with %.3f. */
typedef struct {
	int32_t sread, qread;
	float score;
	int32_t sbeg, send, slen, strand, qbeg, qend, qlen;
} mecat_asm_overlap;

/* replaces load_read + creat_ref_index (:345-372, :397-497): the subject file on the device and its 13-mer lists */
int mecat_b200_asm_index_build(mecat_b200_ctx* ctx, const mecat_asm_reads* subject, void** asmidx);
int mecat_b200_asm_index_release(mecat_b200_ctx* ctx, void* asmidx);
/* replaces pairwise_mapping (:515-984) for the reads of one query file: overlaps with subject reads numbered below the
 * query read, in read order and per read in the order the reference aligns its candidates.  Where the reference reads
 * block memory no seed of the strand wrote, zero is read. */
int mecat_b200_asm_overlaps(mecat_b200_ctx* ctx, void* asmidx, const mecat_asm_reads* query, const mecat_asm_params* p,
                            mecat_asm_overlap** overlaps, size_t* n);
/* test hook: the k-mer lists (begin: 4^13 + 1 entries; positions: 1-based k-mer starts as databaseindex holds them) */
int mecat_b200_asm_index_export(mecat_b200_ctx* ctx, void* asmidx, int64_t* num_positions, uint32_t* begin, int32_t* positions);

/* frees host buffers handed out by this library (same as mecat_b200_free without a context) */
void mecat_b200_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MECAT_B200_H */
