// mecat_b200/csrc/host/mecat2cns.cpp -- host driver with the reference's mecat2cns command line.
//
//   mecat2cns [options] <candidates.can | overlaps.m4> <reads.fasta> <corrected.fasta>
//
// Same flags (src/mecat2cns/options.cpp:201-303), same corrected-FASTA records
// (`>{id}_{beg}_{end}_{len}`, src/mecat2cns/reads_correction_can.cpp:44-46) as the reference's
// `-i 0` path.  The candidate file is normalised in memory exactly like partition_candidates /
// normalise_candidate (src/mecat2cns/overlaps_partition.cpp:141-224) and processed in the same
// batches of `-p` reads, but no `.partN` scratch files are written next to the input.  All gapped
// extensions run on the GPU (mecat_b200_cns_reads_multi); `-t` is accepted and ignored.  The read set may be of any size:
// it is kept as volumes of at most 2.14 Gbase (the splitter's cut, MECAT_VOLUME_BASES as in mecat2pw), all resident on
// every device.
// `-i 1` (M4 input, the default; reads_correction_m4.cpp, overlaps_partition.cpp:345-410) builds the partition records in file
// order; the library orders them like the reference run with one OpenMP thread.  `-x 1` (nanopore) switches the defaults
// and the consensus variant (mecat_correction.cpp:303-360, 453-512).
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <fstream>
#include <thread>
#include <string>
#include <map>
#include <vector>

#include "mecat_b200.h"

namespace {

struct Options
{
	int input_type = 1;
	int num_threads = 1;
	long long batch_size = 100000;
	double min_mapping_ratio = 0.9;
	int min_align_size = 2000;
	int min_cov = 6;
	long long min_size = 5000;
	int tech = 0;
	int num_partition_files = 10;
	bool usage = false;
	const char* overlaps = NULL;
	const char* reads = NULL;
	const char* output = NULL;
};

void print_usage(const char* prog)
{
	fprintf(stderr, "usage:\n%s [options] input reads output\n\noptions:\n", prog);
	fprintf(stderr, "-x <0/1>\tsequencing platform: 0 = PACBIO, 1 = NANOPORE\n\t\tdefault: 0\n");
	fprintf(stderr, "-i <0/1>\tinput type: 0 = candidate, 1 = m4\n");
	fprintf(stderr, "-t <Integer>\tnumber of threads (CPU) -- accepted, unused: extensions run on the GPU\n");
	fprintf(stderr, "-p <Integer>\tbatch size that the reads will be partitioned\n");
	fprintf(stderr, "-r <Real>\tminimum mapping ratio\n");
	fprintf(stderr, "-a <Integer>\tminimum overlap size\n");
	fprintf(stderr, "-c <Integer>\tminimum coverage under consideration\n");
	fprintf(stderr, "-l <Integer>\tminimum length of corrected sequence\n");
	fprintf(stderr, "-k <Integer>\tnumber of partition files when partitioning overlap results\n");
	fprintf(stderr, "-h\t\tprint usage info.\n");
	fprintf(stderr, "\ndefault values (pacbio): -i 1 -t 1 -p 100000 -r 0.9 -a 2000 -c 6 -l 5000 -k 10\n");
	fprintf(stderr, "default values (nanopore, -x 1): -i 1 -t 1 -p 100000 -r 0.4 -a 400 -c 6 -l 2000 -k 10\n");
}

int parse_arguments(int argc, char* argv[], Options& t)
{
	for (int i = 0; i + 1 < argc; ++i)
		if (strcmp(argv[i], "-x") == 0) {
			if (argv[i + 1][0] == '1') t.tech = 1;
			else if (argv[i + 1][0] != '0') { fprintf(stderr, "invalid argument to option 'x': %s\n", argv[i + 1]); return 1; }
			break;
		}
	if (t.tech == 1) { t.min_mapping_ratio = 0.4; t.min_align_size = 400; t.min_size = 2000; }
	int c;
	opterr = 0;
	while ((c = getopt(argc, argv, "i:t:p:r:a:c:l:x:k:h")) != -1) {
		switch (c) {
		case 'i':
			if (optarg[0] == '0') t.input_type = 0;
			else if (optarg[0] == '1') t.input_type = 1;
			else { fprintf(stderr, "invalid argument to option 'i': %s\n", optarg); return 1; }
			break;
		case 't': t.num_threads = atoi(optarg); break;
		case 'p': t.batch_size = atoll(optarg); break;
		case 'r': t.min_mapping_ratio = atof(optarg); break;
		case 'a': t.min_align_size = atoi(optarg); break;
		case 'c': t.min_cov = atoi(optarg); break;
		case 'l': t.min_size = atoll(optarg); break;
		case 'h': t.usage = true; break;
		case 'x': break;
		case 'k': t.num_partition_files = atoi(optarg); break;
		case '?': fprintf(stderr, "unrecognised option '%c'\n", (char)optopt); return 1;
		case ':': fprintf(stderr, "argument to option '%c' is missing.\n", (char)optopt); return 1;
		}
	}
	bool ok = true;
	if (t.num_threads <= 0) { fprintf(stderr, "cpu threads must be greater than 0\n"); ok = false; }
	if (t.batch_size <= 0) { fprintf(stderr, "batch size must be greater than 0\n"); ok = false; }
	if (t.min_mapping_ratio < 0.0) { fprintf(stderr, "mapping ratio must be >= 0.0\n"); ok = false; }
	if (t.min_cov < 0) { fprintf(stderr, "coverage must be >= 0\n"); ok = false; }
	if (argc < 3) return 1;
	t.overlaps = argv[argc - 3];
	t.reads = argv[argc - 2];
	t.output = argv[argc - 1];
	return ok ? 0 : 1;
}

struct StderrTimer
{
	std::string name;
	timeval t0;
	explicit StderrTimer(const std::string& n) : name(n) { fprintf(stderr, "[%s] begins.\n", name.c_str()); gettimeofday(&t0, NULL); }
	~StderrTimer()
	{
		timeval t1;
		gettimeofday(&t1, NULL);
		fprintf(stderr, "[%s] takes %.2f secs.\n", name.c_str(), t1.tv_sec - t0.tv_sec + 1.0 * (t1.tv_usec - t0.tv_usec) / 1000000);
	}
};

// `.can` lines: qid sid qdir sdir qext sext score qsize ssize (src/common/alignment.cpp:9-16)
bool load_candidates(const char* path, std::vector<mecat_candidate>& out)
{
	FILE* f = fopen(path, "rb");
	if (!f) return false;
	// the whole file at once, then nine decimal integers per line (a million lines through sscanf took 0.4 s)
	std::string buf;
	{
		char chunk[1 << 16];
		size_t n;
		while ((n = fread(chunk, 1, sizeof chunk, f)) > 0) buf.append(chunk, n);
	}
	fclose(f);
	const char* p = buf.data();
	const char* end = p + buf.size();
	while (p < end) {
		const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
		if (!eol) eol = end;
		int v[9], k = 0;
		const char* q = p;
		while (k < 9) {
			while (q < eol && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
			if (q >= eol) break;
			bool neg = false;
			if (*q == '-' || *q == '+') { neg = *q == '-'; ++q; }
			if (q >= eol || *q < '0' || *q > '9') break;
			long long x = 0;
			while (q < eol && *q >= '0' && *q <= '9') { x = x * 10 + (*q - '0'); ++q; }
			v[k++] = (int)(neg ? -x : x);
		}
		if (k == 9) {
			mecat_candidate e;
			memset(&e, 0, sizeof e);
			e.qid = v[0]; e.sid = v[1]; e.qdir = v[2]; e.sdir = v[3]; e.qext = v[4]; e.sext = v[5]; e.score = v[6]; e.qsize = v[7]; e.ssize = v[8];
			out.push_back(e);
		}
		p = eol + 1;
	}
	return true;
}

// `.m4` lines of mecat2pw -j 1 -g 1 (operator>>(M4Record), src/common/alignment.cpp:34-56) -> the records partition_m4records
// writes (src/mecat2cns/overlaps_partition.cpp:345-410): size and mapping-range filters, then m4_to_candidate of
// normalize_m4record for either read as the one to correct (src/common/alignment.h:71-103,170-186), in file order, into
// the partition of that read (batch_size reads each).  0 = ok, 1 = cannot open, 2 = no extension points in the file.
int load_m4_partitions(const char* path, double min_cov_ratio, long long batch_size, long long min_read_size,
                       std::map<long long, std::vector<mecat_candidate>>& parts)
{
	FILE* f = fopen(path, "r");
	if (!f) return 1;
	char line[1024];
	while (fgets(line, sizeof line, f)) {
		long long qid, sid, qoff, qend, qsize, soff, send, ssize, qext = -1, sext = -1;
		double ident;
		int vscore, qdir, sdir;
		const int n = sscanf(line, "%lld %lld %lf %d %d %lld %lld %lld %d %lld %lld %lld %lld %lld", &qid, &sid, &ident, &vscore, &qdir, &qoff, &qend,
		                     &qsize, &sdir, &soff, &send, &ssize, &qext, &sext);
		if (n < 12) continue;
		if (n < 14) { fclose(f); return 2; }
		if (qsize < min_read_size || ssize < min_read_size) continue;
		const long long qm = qend - qoff, qs = (long long)((double)qsize * min_cov_ratio), sm = send - soff, ss = (long long)((double)ssize * min_cov_ratio);
		if (!(qm >= qs || sm >= ss)) continue;            // check_m4record_mapping_range
		for (int subject_is_target = 0; subject_is_target < 2; ++subject_is_target) {
			mecat_candidate e;
			memset(&e, 0, sizeof e);
			if (subject_is_target) {
				e.qdir = qdir; e.qid = (int32_t)qid; e.qext = (int32_t)qext; e.qsize = (int32_t)qsize; e.qoff = (int32_t)qoff; e.qend = (int32_t)qend;
				e.sdir = sdir; e.sid = (int32_t)sid; e.sext = (int32_t)sext; e.ssize = (int32_t)ssize; e.soff = (int32_t)soff; e.send = (int32_t)send;
			} else {                                      // reverse_m4record
				e.qdir = sdir; e.qid = (int32_t)sid; e.qext = (int32_t)sext; e.qsize = (int32_t)ssize; e.qoff = (int32_t)soff; e.qend = (int32_t)send;
				e.sdir = qdir; e.sid = (int32_t)qid; e.sext = (int32_t)qext; e.ssize = (int32_t)qsize; e.soff = (int32_t)qoff; e.send = (int32_t)qend;
			}
			e.score = vscore;
			if (e.sdir == 1) { e.sdir = 0; e.qdir = 1 - e.qdir; }
			parts[e.sid / batch_size].push_back(e);
		}
	}
	fclose(f);
	return 0;
}

// normalise_candidate, overlaps_partition.cpp:141-165
mecat_candidate normalise(const mecat_candidate& s, bool subject_is_target)
{
	mecat_candidate d;
	memset(&d, 0, sizeof d);
	if (subject_is_target) {
		d.qdir = s.qdir; d.qid = s.qid; d.qext = s.qext; d.qsize = s.qsize;
		d.sdir = s.sdir; d.sid = s.sid; d.sext = s.sext; d.ssize = s.ssize;
	} else {
		d.qdir = s.sdir; d.qid = s.sid; d.qext = s.sext; d.qsize = s.ssize;
		d.sdir = s.qdir; d.sid = s.qid; d.sext = s.qext; d.ssize = s.qsize;
	}
	d.score = s.score;
	if (d.sdir == 1) { d.qdir = 1 - d.qdir; d.sdir = 1 - d.sdir; }
	return d;
}

}  // namespace

int main(int argc, char* argv[])
{
	Options opt;
	const int r = parse_arguments(argc, argv, opt);
	if (r) { print_usage(argv[0]); return 1; }
	if (opt.usage) { print_usage(argv[0]); return 0; }
	if (mecat_b200_device_count() < 1) { fprintf(stderr, "mecat2cns: no CUDA device found (this build has no CPU path)\n"); return 1; }

	// Creating the CUDA context of device 0 takes about a second; it runs next to the candidate and FASTA loading.
	mecat_b200_ctx* ctx0 = NULL;
	int ctx0_rc = 0;
	std::thread warm([&]() { ctx0_rc = mecat_b200_init(&ctx0, 0, NULL); });
	struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } warm_joiner{warm};
	std::vector<mecat_candidate> raw, ec;
	std::map<long long, std::vector<mecat_candidate>> m4_parts;      // -i 1: the partitions, records in file order
	if (opt.input_type == 1) {
		StderrTimer t("partition_m4records");
		const int rc = load_m4_partitions(opt.overlaps, opt.min_mapping_ratio - 0.02, opt.batch_size, opt.min_size, m4_parts);
		if (rc == 1) { fprintf(stderr, "cannot open file '%s' for reading\n", opt.overlaps); return 1; }
		if (rc == 2) { fprintf(stderr, "no gapped start position is provided, please make sure that you have run 'meap_pairwise' with option '-g 1'\n"); return 1; }
		// one call per partition (the reference orders a partition as a whole); the partitions one after the other in `ec`
		for (auto& kv : m4_parts) ec.insert(ec.end(), kv.second.begin(), kv.second.end());
	} else {
		StderrTimer t("partition_candidates");
		if (!load_candidates(opt.overlaps, raw)) { fprintf(stderr, "cannot open file '%s' for reading\n", opt.overlaps); return 1; }
		ec.reserve(raw.size() * 2);
		for (const mecat_candidate& e : raw) {
			if (e.qsize < opt.min_size || e.ssize < opt.min_size) continue;     // overlaps_partition.cpp:205
			ec.push_back(normalise(e, false));
			ec.push_back(normalise(e, true));
		}
		raw.clear(); raw.shrink_to_fit();
	}

	// all reads, packed, as one or more volumes (the reference keeps them in one PackedDB with 64-bit offsets, packed_db.cpp:194)
	mecat_volume* vols = NULL;
	int nvols = 0;
	{
		StderrTimer t("load_fasta_db");
		char err[512];
		const char* cap = getenv("MECAT_VOLUME_BASES");   // test hook: smaller volumes, as in mecat2pw
		if (mecat_b200_volumes_from_fasta(opt.reads, cap ? atoll(cap) : 0, &vols, &nvols, err, sizeof err)) { fprintf(stderr, "mecat2cns: %s\n", err); return 1; }
	}

	// One host thread per GPU (MECAT_GPUS=n, default 1): every device holds a replica of the packed reads, the reads to
	// correct are cut into contiguous slices of the id-sorted candidate list (balanced by candidate count, cut at read
	// boundaries), no data moves between devices.  Slices are written in id order, so the output does not depend on n.
	int ngpus = 1;
	if (const char* g = getenv("MECAT_GPUS")) ngpus = atoi(g);
	const int have = mecat_b200_device_count();
	if (ngpus < 1) ngpus = 1;
	if (ngpus > have) ngpus = have;
	if (opt.input_type == 0) std::stable_sort(ec.begin(), ec.end(), [](const mecat_candidate& a, const mecat_candidate& b) { return a.sid < b.sid; });
	const mecat_cns_params P = {opt.min_mapping_ratio, opt.min_align_size, opt.min_cov, opt.min_size, opt.tech, opt.input_type};
	std::vector<size_t> cut((size_t)ngpus + 1, ec.size());
	cut[0] = 0;
	for (int g = 1; g < ngpus; ++g) {
		size_t k = ec.size() * (size_t)g / (size_t)ngpus;
		// -i 0: cut at a read boundary; -i 1: at a partition boundary (a partition is ordered as a whole)
		if (opt.input_type == 0) while (k > 0 && k < ec.size() && ec[k].sid == ec[k - 1].sid) ++k;
		else while (k > 0 && k < ec.size() && ec[k].sid / opt.batch_size == ec[k - 1].sid / opt.batch_size) ++k;
		cut[g] = std::max(k, cut[g - 1]);
	}
	if (warm.joinable()) warm.join();
	if (ctx0_rc) ctx0 = NULL;
	// MECAT_B200_FAST_EXIT (default 1): skip the block-by-block release of the devices' memory pools at the end; the
	// process exits as soon as the corrected reads are written and closed (0: explicit release, as a library user would)
	const char* fe = getenv("MECAT_B200_FAST_EXIT");
	const bool fast_exit = !fe || atoi(fe) != 0;
	struct Part { long long part; std::string text; };
	std::vector<std::vector<Part>> results((size_t)ngpus);
	std::atomic<int> failed(0);
	auto worker = [&](int dev) {
		mecat_b200_ctx* ctx = dev == 0 ? ctx0 : NULL;
		std::vector<void*> dvols((size_t)nvols, (void*)NULL);
		{
			StderrTimer t("gpu " + std::to_string(dev) + " init + volume upload");
			if (!ctx && mecat_b200_init(&ctx, dev, NULL)) { fprintf(stderr, "mecat2cns: cannot initialise GPU %d\n", dev); failed = 1; return; }
			for (int v = 0; v < nvols; ++v)
				if (mecat_b200_volume_upload(ctx, &vols[v], &dvols[(size_t)v])) { fprintf(stderr, "mecat2cns: %s\n", mecat_b200_last_error(ctx)); failed = 1; return; }
		}
		for (size_t i = cut[dev]; !failed && i < cut[dev + 1];) {
			const long long part = ec[i].sid / opt.batch_size;
			size_t j = i;
			while (j < cut[dev + 1] && ec[j].sid / opt.batch_size == part) ++j;
			char info[128];
			snprintf(info, sizeof info, "gpu %d processing reads %lld --- %lld", dev, part * opt.batch_size, (part + 1) * opt.batch_size - 1);
			StderrTimer t(info);
			mecat_cns_piece* pieces = NULL;
			char* seqs = NULL;
			size_t np = 0, nb = 0;
			if (nvols == 0) { i = j; continue; }
			if (mecat_b200_cns_reads_multi(ctx, dvols.data(), nvols, ec.data() + i, j - i, &P, &pieces, &np, &seqs, &nb)) {
				fprintf(stderr, "mecat2cns: %s\n", mecat_b200_last_error(ctx));
				failed = 1;
				break;
			}
			Part out;
			out.part = part;
			out.text.reserve(nb + 48 * np);
			char head[128];
			for (size_t k = 0; k < np; ++k) {
				const int hl = snprintf(head, sizeof head, ">%lld_%lld_%lld_%lld\n", (long long)pieces[k].id, (long long)pieces[k].beg,
				                        (long long)pieces[k].end, (long long)pieces[k].seq_len);
				out.text.append(head, (size_t)hl);
				out.text.append(seqs + pieces[k].seq_offset, (size_t)pieces[k].seq_len);
				out.text.push_back('\n');
			}
			results[(size_t)dev].push_back(std::move(out));
			mecat_b200_free(ctx, pieces);
			mecat_b200_free(ctx, seqs);
			i = j;
		}
		if (dev == 0 && getenv("MECAT_B200_STATS")) {       // per-kernel CUDA-event times of device 0's share, one line
			mecat_b200_stats st;
			if (!mecat_b200_get_stats(ctx, &st)) {
				static const char* names[MECAT_K_NUM] = {"orient", "index_count", "scan", "index_fill", "index_sort", "seed", "walk", "merge", "extend",
				                                         "finalize", "cns_accept", "cns_normvote", "cns_segment", "cns_region", "cns_poa", "cns_assemble",
				                                         "ref_count", "ref_seed", "ref_rescue", "asm_index", "asm_seed", "asm_extend"};
				fprintf(stderr, "[kernel ms]");
				for (int k = 0; k < MECAT_K_NUM; ++k)
					if (st.kernel_launches[k]) fprintf(stderr, " %s=%.1f(%lld)", names[k], st.kernel_ms[k], (long long)st.kernel_launches[k]);
				fprintf(stderr, "\n");
			}
		}
		if (fast_exit) return;                 // the process ends right after the output is written: nothing to hand back
		{
			StderrTimer t("gpu " + std::to_string(dev) + " release");
			for (void* d : dvols) if (d) mecat_b200_volume_release(ctx, d);
			mecat_b200_destroy(ctx);
		}
	};
	{
		std::vector<std::thread> th;
		for (int d = 1; d < ngpus; ++d) th.emplace_back(worker, d);
		worker(0);
		for (auto& t : th) t.join();
	}
	const bool ok = !failed;
	if (ok) {
		StderrTimer t("write results");
		std::ofstream out(opt.output, std::ios::binary);
		if (!out) { fprintf(stderr, "cannot open '%s' for writing\n", opt.output); return 1; }
		for (auto& per_dev : results) for (auto& pt : per_dev) out.write(pt.text.data(), (std::streamsize)pt.text.size());
		out.close();
		if (!out) { fprintf(stderr, "cannot write '%s'\n", opt.output); return 1; }
	}
	if (ok && fast_exit) { fflush(stdout); fflush(stderr); _exit(0); }
	mecat_b200_volumes_unload(vols, nvols);
	return ok ? 0 : 1;
}
