// mecat_b200/csrc/host/mecat2ref.cpp -- host driver with the reference's mecat2ref command line.
//
//   mecat2ref -d reads -r reference -o output -w wrk_dir [-t threads] [-n candidates] [-b best] [-m 0|1] [-x 0]
//
// Same flags and defaults as src/mecat2ref/mecat2ref.cpp:53-190, same records as its ref (-m 0), m4 (-m 1) and sam (-m 2)
// output (src/mecat2ref/output.cpp:8-191; the records of a read are together, reads in input order).  The genome is indexed once
// on the GPU (mecat_b200_ref_index_build); the reads go through mecat_b200_ref_map in batches -- seeding, DDF scoring,
// gapped extension and clipped-end rescue all run on the device.  `-t` only sizes the host threads that pack the reads.  With MECAT_GPUS=n every
// device holds a replica of the genome index and maps its share of the read batches (no collective; the output does not
// depend on n).  Not on this path (refused with a message): -x 1 (nanopore).  The reference's scratch
// files (wrk_dir/N.fq, N.r, chrindex.txt, ./config.txt) are not written; the working directory is still created.
#include <dirent.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>

#include <algorithm>
#include <atomic>
#include <future>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "mecat_b200.h"
#include "refio.h"

namespace {

struct Options
{
	const char* reads = NULL;
	const char* reference = NULL;
	const char* wrk_dir = NULL;
	const char* output = NULL;
	int num_cores = 1, num_candidates = 10, num_output = 10, output_format = 0, tech = 0;
};

void print_usage(const char* prog)
{
	fprintf(stderr, "\n\nusage:\n%s [-d reads] [-r reference] [-o output] [-w working dir] [-t threads]\n\noptions:\n", prog);
	fprintf(stderr, "-d <string>\treads file name\n-r <string>\treference file name\n-o <string>\toutput file name\n");
	fprintf(stderr, "-w <string>\tworking folder name, will be created if not exist\n");
	fprintf(stderr, "-t <integer>\tnumber of cput threads (host-side packing only: the mapping runs on the GPU)\n\t\tdefault: 1\n");
	fprintf(stderr, "-n <integer>\tnumber of of candidates for gap extension\n\t\tdefault: 10\n");
	fprintf(stderr, "-b <integer>\toutput the best b alignments\n\t\tdefault: 10\n");
	fprintf(stderr, "-m <0/1/2>\toutput format: 0 = ref, 1 = m4, 2 = sam\n\t\tdefault: 0\n");
	fprintf(stderr, "-x <0/1>\tsequencing technology: 0 = pacbio, 1 = nanopore\n\t\tdefault: 0\n");
}

int parse(int argc, char* argv[], Options& o)
{
	int c;
	opterr = 0;
	while ((c = getopt(argc, argv, "d:r:w:o:t:n:b:m:x:")) != -1) {
		switch (c) {
		case 'd': o.reads = optarg; break;
		case 'r': o.reference = optarg; break;
		case 'w': o.wrk_dir = optarg; break;
		case 'o': o.output = optarg; break;
		case 't': o.num_cores = atoi(optarg); break;
		case 'n': o.num_candidates = atoi(optarg); break;
		case 'b': o.num_output = atoi(optarg); break;
		case 'm': o.output_format = atoi(optarg); break;
		case 'x':
			if (optarg[0] == '0') o.tech = 0;
			else if (optarg[0] == '1') o.tech = 1;
			else { fprintf(stderr, "invalid argument to option 'x': %s\n", optarg); return -1; }
			break;
		default:
			if (optopt && strchr("drwotnbmx", optopt)) fprintf(stderr, "Error: argument to option '%c' is missing!\n", optopt);
			else fprintf(stderr, "Error: unrecogised option '%c'\n", optopt);
			return -1;
		}
	}
	const char* msg = NULL;
	if (!o.reads) msg = "dataset must be specified";
	else if (!o.reference) msg = "reference must be specified";
	else if (!o.output) msg = "output must be specified";
	else if (!o.wrk_dir) msg = "working directory must be specified";
	else if (o.num_cores < 1) msg = "cpu cores must be > 0";
	else if (o.num_candidates < 1) msg = "candidates must be > 0";
	else if (o.num_output < 1) msg = "output alignments must be > 0";
	if (msg) { fprintf(stderr, "Error: %s\n", msg); return -1; }
	if (o.num_output > o.num_candidates) {
		fprintf(stderr, "warning: number of output (%d) is greater than number of candidates (%d), we reset it to %d", o.num_output, o.num_candidates,
		        o.num_candidates);
		o.num_output = o.num_candidates;
	}
	DIR* d = opendir(o.wrk_dir);
	if (d) closedir(d);
	else if (mkdir(o.wrk_dir, S_IRWXU) == -1) { fprintf(stderr, "Fail to create folder %s!\n", o.wrk_dir); return -1; }
	return 0;
}

double now()
{
	struct timeval t;
	gettimeofday(&t, NULL);
	return (double)t.tv_sec + 1e-6 * (double)t.tv_usec;
}

struct Batch { int first, count; std::string text; };

}  // namespace

int main(int argc, char* argv[])
{
	Options o;
	if (parse(argc, argv, o) == -1) { print_usage(argv[0]); return 1; }

	if (o.output_format < 0 || o.output_format > 2) { fprintf(stderr, "mecat2ref (b200): unknown output format %d (0 = ref, 1 = m4, 2 = sam)\n", o.output_format); return 1; }
	const double t0 = now();
	{
		// before any work: an unwritable output should not cost a mapping run -- but an existing result is only replaced once
		// the new one is complete (append mode creates, never truncates)
		FILE* probe = fopen(o.output, "a");
		if (!probe) { fprintf(stderr, "failed to open file %s for writing.\n", o.output); return 1; }
		fclose(probe);
	}
	int ndev = 1;
	if (const char* e = getenv("MECAT_GPUS")) ndev = std::max(1, atoi(e));
	if (ndev > mecat_b200_device_count()) { fprintf(stderr, "mecat2ref (b200): MECAT_GPUS=%d but %d CUDA device(s) visible\n", ndev, mecat_b200_device_count()); return 1; }

	// the CUDA contexts come up while the host parses and packs the genome
	std::vector<mecat_b200_ctx*> ctx((size_t)ndev, (mecat_b200_ctx*)NULL);
	std::vector<int> init_rc((size_t)ndev, 0);
	std::vector<std::thread> starters;
	for (int d = 0; d < ndev; ++d) starters.emplace_back([&, d]() { init_rc[(size_t)d] = mecat_b200_init(&ctx[(size_t)d], d, NULL); });
	refio::Genome G;
	refio::Reads R;
	std::string err;
	const bool loaded = refio::load_genome(o.reference, G, err) && refio::load_reads(o.reads, R, err);
	for (auto& t : starters) t.join();
	for (int d = 0; d < ndev; ++d)
		if (init_rc[(size_t)d]) { fprintf(stderr, "mecat2ref (b200): no usable CUDA device %d (there is no CPU fallback)\n", d); return 1; }
	if (!loaded) { fprintf(stderr, "mecat2ref (b200): %s\n", err.c_str()); return 1; }
	const double t_load = now();

	// batches of reads: one ABI call each.  A volume holds < 2^31 bases; the ref format returns two strings per record.
	int64_t max_bases = 1000000000ll;            // explicit reverse strands of reads with other letters still fit
	if (const char* e = getenv("MECAT_B200_REF_BATCH_BASES")) max_bases = std::max(1ll, atoll(e));      // test hook: many small batches
	const int max_reads = o.output_format != 1 ? 20000 : 1 << 30;
	std::vector<Batch> batches;
	const int total = (int)R.size();
	const int64_t all_bases = R.total + total;
	const int64_t share = std::max<int64_t>(std::min<int64_t>(1 << 20, max_bases), (all_bases + ndev - 1) / ndev);
	for (int first = 0; first < total;) {
		int count = 0;
		int64_t bases = 0;
		while (first + count < total && count < max_reads) {
			const int64_t need = R.length((size_t)(first + count)) + 1;
			if (count && (bases + need > max_bases || bases + need > share)) break;
			bases += need; ++count;
		}
		Batch b; b.first = first; b.count = count;
		batches.push_back(b);
		first += count;
	}

	const int hw = (int)std::thread::hardware_concurrency();
	const int host_threads = std::max(1, std::min(hw > 0 ? hw : 1, std::max(o.num_cores, 8)) / ndev);
	std::atomic<int> failed(0);
	std::vector<double> t_index((size_t)ndev, 0.0);
	auto pack = [&](int k) {
		std::unique_ptr<refio::ReadBatch> B(new refio::ReadBatch);
		B->build(R, (size_t)batches[(size_t)k].first, (size_t)batches[(size_t)k].count, host_threads);
		return B;
	};
	// Device d takes batches d, d + ndev, ...  While one batch is on the device the host packs the next one and writes
	// the text of the previous one.
	auto worker = [&](int d) {
		mecat_b200_ctx* c = ctx[(size_t)d];
		std::future<std::unique_ptr<refio::ReadBatch>> packed;
		std::future<void> text;
		if (d < (int)batches.size()) packed = std::async(std::launch::async, pack, d);
		const double a = now();
		void* idx = NULL;
		const mecat_ref_genome g = G.view();
		if (mecat_b200_ref_index_build(c, &g, &idx)) { fprintf(stderr, "mecat2ref (b200): %s\n", mecat_b200_last_error(c)); failed = 1; }
		t_index[(size_t)d] = now() - a;
		mecat_ref_params p;
		p.num_candidates = o.num_candidates; p.num_output = o.num_output; p.want_strings = o.output_format != 1; p.tech = o.tech;
		for (int k = d; k < (int)batches.size(); k += ndev) {
			std::unique_ptr<refio::ReadBatch> B = packed.get();
			if (k + ndev < (int)batches.size()) packed = std::async(std::launch::async, pack, k + ndev);
			if (failed) continue;                   // keep draining the packer
			const mecat_ref_reads view = B->view();
			mecat_ref_result* res = NULL;
			char *qs = NULL, *ss = NULL;
			size_t n = 0, nbytes = 0;
			if (mecat_b200_ref_map(c, idx, &view, &p, &res, &n, &qs, &ss, &nbytes)) { fprintf(stderr, "mecat2ref (b200): %s\n", mecat_b200_last_error(c)); failed = 1; continue; }
			if (text.valid()) text.get();
			Batch* b = &batches[(size_t)k];
			text = std::async(std::launch::async, [&, b, res, n, qs, ss]() {
				refio::format_results(b->text, G, R.name, b->first, res, n, qs, ss, o.output_format);
				mecat_b200_host_free(res); mecat_b200_host_free(qs); mecat_b200_host_free(ss);
			});
		}
		if (text.valid()) text.get();
		if (!idx) return;
		if (d == 0 && getenv("MECAT_B200_STATS")) {       // per-kernel CUDA-event times of device 0's share, one line
			mecat_b200_stats st;
			if (!mecat_b200_get_stats(c, &st)) {
				static const char* names[MECAT_K_NUM] = {"orient", "index_count", "scan", "index_fill", "index_sort", "seed", "walk", "merge", "extend",
				                                         "finalize", "cns_accept", "cns_normvote", "cns_segment", "cns_region", "cns_poa", "cns_assemble",
				                                         "ref_count", "ref_seed", "ref_rescue", "asm_index", "asm_seed", "asm_extend"};
				fprintf(stderr, "[kernel ms]");
				for (int k = 0; k < MECAT_K_NUM; ++k)
					if (st.kernel_launches[k]) fprintf(stderr, " %s=%.1f(%lld)", names[k], st.kernel_ms[k], (long long)st.kernel_launches[k]);
				fprintf(stderr, " hits=%lld extensions=%lld records=%lld h2d=%.1fMB d2h=%.1fMB\n", (long long)st.num_hits, (long long)st.num_candidates,
				        (long long)st.num_records, st.h2d_bytes / 1e6, st.d2h_bytes / 1e6);
			}
		}
		mecat_b200_ref_index_release(c, idx);
	};
	std::vector<std::thread> workers;
	for (int d = 0; d < ndev; ++d) workers.emplace_back(worker, d);
	for (auto& t : workers) t.join();
	if (failed) return 1;
	const double t_map = now();

	fprintf(stderr, "output file name: %s\n", o.output);
	FILE* out = fopen(o.output, "w");
	if (!out) { fprintf(stderr, "failed to open file %s for writing.\n", o.output); return 1; }
	if (o.output_format == 2) {
		std::string head;
		refio::sam_header(head, G, argc, argv);
		fwrite(head.data(), 1, head.size(), out);
	}
	bool wrote = true;
	for (const Batch& b : batches) wrote = fwrite(b.text.data(), 1, b.text.size(), out) == b.text.size() && wrote;
	wrote = (fclose(out) == 0) && wrote;
	if (!wrote) { fprintf(stderr, "mecat2ref (b200): cannot write %s\n", o.output); return 1; }
	// MECAT_B200_FAST_EXIT (default 1): the output is on disk; releasing the devices' memory pools block by block would only
	// delay the exit (0: explicit release)
	const char* fe = getenv("MECAT_B200_FAST_EXIT");
	const bool fast_exit = !fe || atoi(fe) != 0;
	if (!fast_exit) for (int d = 0; d < ndev; ++d) mecat_b200_destroy(ctx[(size_t)d]);
	const double t1 = now();
	fprintf(stderr, "mecat2ref (b200): %d reads, %lld reference bases, %d device(s): load %.2f s, index %.2f s, mapping %.2f s, write %.2f s, total %.2f s\n",
	        total, (long long)G.seq.n, ndev, t_load - t0, t_index[0], t_map - t_load - t_index[0], t1 - t_map, t1 - t0);
	if (fast_exit) { fflush(stdout); fflush(stderr); _exit(0); }
	return 0;
}
