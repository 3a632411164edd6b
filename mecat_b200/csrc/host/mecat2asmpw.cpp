// mecat_b200/csrc/host/mecat2asmpw.cpp -- host driver with the command line of mecat2canu's corrected-read overlappers:
//
//   mecat2asmpw -P<blocks dir> -T<threads> -S<first file> -E<last file>
//
// (mecat2canu/src/mecat2asmpw/mecat2asmpw.c:1032-1166; the pipeline's call is Overlapmecat2asmpw.pm:483-496).  One source,
// four programs: the name decides the variant -- `mecat2trimpw*` uses the trimming gate and score (mecat2trimpw.c:640,
// :942-943), a name ending in `50` keeps 50 candidates per read instead of 100 (MAXC, mecat2asmpw50.c:23).
// `<dir>/ovlprep` names the read range of every block file (`-allreads -allbases -b <first> -e <last>` per line, :1083-1090),
// `<dir>/NNNNNN.fasta` hold the reads (header line, one sequence line).  The index is built over file S; the reads of the
// files S .. E are mapped against it (:1112-1150) on the GPU (mecat_b200_asm_index_build / mecat_b200_asm_overlaps); `-T`
// is accepted and only decides how many `<S>_<t>.r` result files exist: the reference writes one per thread and the
// pipeline concatenates `<S>_*.r`; here device k writes `<S>_<k mod T>.r` and the others stay empty.
// Several devices (MECAT_GPUS=n, default 1; MECAT_DEVICE names the first): every device holds its own index of file S and
// maps its share of every query file's reads (contiguous slices) -- replicas sharded by read, no exchange.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mecat_b200.h"
#include "format.h"

namespace {

struct File { std::string text; std::vector<int64_t> start; std::vector<int32_t> len; };

// load_read / load_fastq (:345-372, :1000-1029): a header line, then the sequence as one token
bool load_fasta(const std::string& path, File& f)
{
	FILE* in = fopen(path.c_str(), "rb");
	if (!in) return false;
	std::string buf;
	char tmp[1 << 16];
	size_t n;
	while ((n = fread(tmp, 1, sizeof tmp, in)) > 0) buf.append(tmp, n);
	fclose(in);
	f.text.clear(); f.start.clear(); f.len.clear();
	f.text.reserve(buf.size());
	size_t p = 0;
	while (p < buf.size()) {
		size_t e = buf.find('\n', p);
		if (e == std::string::npos) e = buf.size();
		if (buf[p] == '>') {
			size_t q = e + 1;
			while (q < buf.size() && (buf[q] == '\n' || buf[q] == '\r' || buf[q] == ' ' || buf[q] == '\t')) ++q;
			size_t r = q;
			while (r < buf.size() && buf[r] != '\n' && buf[r] != '\r' && buf[r] != ' ' && buf[r] != '\t') ++r;
			if (r > q && buf[q] != '>') {
				f.start.push_back((int64_t)f.text.size());
				f.len.push_back((int32_t)(r - q));
				f.text.append(buf, q, r - q);
				f.text.push_back('\0');
				e = r;
			}
		}
		p = e + 1;
	}
	return true;
}

double now()
{
	timeval t;
	gettimeofday(&t, NULL);
	return t.tv_sec + 1e-6 * t.tv_usec;
}

}  // namespace

int main(int argc, char** argv)
{
	const char* base = strrchr(argv[0], '/');
	base = base ? base + 1 : argv[0];
	mecat_asm_params P;
	P.variant = strstr(base, "trimpw") ? 1 : 0;
	const size_t bl = strlen(base);
	P.max_candidates = bl >= 2 && !strcmp(base + bl - 2, "50") ? 50 : 100;
	std::string dir;
	int threads = -1, first = -1, last = -1;
	for (int i = 1; i < argc; ++i) {           // param_read :1032-1061: the value follows the letter without a space
		if (argv[i][0] != '-') { fprintf(stderr, "%s: unexpected argument '%s'\n", base, argv[i]); return 1; }
		const char* v = argv[i] + 2;
		switch (argv[i][1]) {
		case 'P': dir = v; break;
		case 'T': threads = atoi(v); break;
		case 'S': first = atoi(v); break;
		case 'E': last = atoi(v); break;
		}
	}
	if (dir.empty() || threads < 1 || first < 1 || last < first) {
		fprintf(stderr, "usage: %s -P<blocks dir> -T<threads> -S<first file> -E<last file>\n", base);
		return 1;
	}
	std::vector<int> file_first, file_last;
	{
		FILE* fp = fopen((dir + "/ovlprep").c_str(), "r");
		if (!fp) { fprintf(stderr, "%s: cannot open %s/ovlprep\n", base, dir.c_str()); return 1; }
		char a[300], b[300], c[300], d[300];
		int lo, hi;
		while ((int)file_first.size() < last && fscanf(fp, " %299s %299s %299s %d %299s %d", a, b, c, &lo, d, &hi) == 6) { file_first.push_back(lo); file_last.push_back(hi); }
		fclose(fp);
	}
	if ((int)file_first.size() < last) { fprintf(stderr, "%s: ovlprep names %zu files, -E asks for %d\n", base, file_first.size(), last); return 1; }
	auto block_path = [&](int i) { char n[32]; snprintf(n, sizeof n, "/%06d.fasta", i); return dir + n; };

	const double t0 = now();
	File sub;
	if (!load_fasta(block_path(first), sub) || sub.len.empty()) { fprintf(stderr, "%s: no reads in %s\n", base, block_path(first).c_str()); return 1; }
	if ((int)sub.len.size() != file_last[first - 1] - file_first[first - 1] + 1) {
		fprintf(stderr, "%s: %s holds %zu reads, ovlprep says %d\n", base, block_path(first).c_str(), sub.len.size(), file_last[first - 1] - file_first[first - 1] + 1);
		return 1;
	}
	if (sub.text.size() >= 0x7fffffffu - 4000u) { fprintf(stderr, "%s: %s holds %zu letters; the program keeps positions in 32 bits\n", base, block_path(first).c_str(), sub.text.size()); return 1; }
	std::vector<int32_t> sub_start(sub.start.begin(), sub.start.end());
	const int64_t part_reads = getenv("MECAT_B200_ASM_PART_READS") ? atoll(getenv("MECAT_B200_ASM_PART_READS")) : 200000;      // test hook: small parts
	const char* dev = getenv("MECAT_DEVICE");
	const int dev0 = dev ? atoi(dev) : 0;
	int ngpu = getenv("MECAT_GPUS") ? atoi(getenv("MECAT_GPUS")) : 1;
	if (ngpu < 1) ngpu = 1;
	if (dev0 + ngpu > mecat_b200_device_count()) {
		fprintf(stderr, "%s: devices %d .. %d asked for, %d present (this program has no CPU path)\n", base, dev0, dev0 + ngpu - 1, mecat_b200_device_count());
		return 1;
	}
	std::vector<FILE*> out((size_t)threads, (FILE*)NULL);
	std::vector<std::mutex> out_mu((size_t)threads);
	for (int t = 0; t < threads; ++t) {
		char n[64];
		snprintf(n, sizeof n, "/%d_%d.r", first, t);
		out[(size_t)t] = fopen((dir + n).c_str(), "w");
		if (!out[(size_t)t]) { fprintf(stderr, "%s: cannot write %s%s\n", base, dir.c_str(), n); return 1; }
	}
	// the query files, loaded once and shared by the devices
	std::vector<File> qfiles((size_t)(last - first + 1));
	for (int i = first + 1; i <= last; ++i)
		if (!load_fasta(block_path(i), qfiles[(size_t)(i - first)])) { fprintf(stderr, "%s: cannot read %s\n", base, block_path(i).c_str()); return 1; }
	const double t_init = now();
	std::vector<double> t_dev((size_t)ngpu, 0.0), t_index((size_t)ngpu, 0.0), t_map((size_t)ngpu, 0.0), t_write((size_t)ngpu, 0.0);
	std::vector<size_t> totals((size_t)ngpu, 0);
	std::vector<int> rcs((size_t)ngpu, 0);
	// MECAT_B200_FAST_EXIT (default 1): once the result files are closed the process ends; handing the devices' memory pools
	// back block by block would only delay that (0: explicit release)
	const char* fe = getenv("MECAT_B200_FAST_EXIT");
	const bool fast_exit = !fe || atoi(fe) != 0;
	auto work = [&](int k) {
		mecat_b200_ctx* ctx = NULL;
		const double d0 = now();
		if (mecat_b200_init(&ctx, dev0 + k, NULL)) { fprintf(stderr, "%s: no CUDA device %d (this program has no CPU path)\n", base, dev0 + k); rcs[(size_t)k] = 1; return; }
		const double i0 = now();
		t_dev[(size_t)k] = i0 - d0;
		mecat_asm_reads S;
		S.text = sub.text.data(); S.num_letters = (int64_t)sub.text.size(); S.num_reads = (int32_t)sub.len.size(); S.first_read_id = file_first[first - 1];
		S.read_start = sub_start.data(); S.read_len = sub.len.data();
		void* idx = NULL;
		if (mecat_b200_asm_index_build(ctx, &S, &idx)) { fprintf(stderr, "%s: %s\n", base, mecat_b200_last_error(ctx)); rcs[(size_t)k] = 1; mecat_b200_destroy(ctx); return; }
		t_index[(size_t)k] = now() - i0;
		const int slot = k % threads;
		for (int i = first; i <= last && !rcs[(size_t)k]; ++i) {
			const File& q = i == first ? sub : qfiles[(size_t)(i - first)];
			// a query file is taken in parts like load_fastq does (:1000-1029: at most SVM = 200 000 reads and MAXSTR = 10^9
			// letters at a time), every part split between the devices by reads
			const int64_t nq = (int64_t)q.len.size();
			for (int64_t a = 0; a < nq && !rcs[(size_t)k];) {
				int64_t b = a;
				while (b < nq && b - a < part_reads && q.start[(size_t)b] + q.len[(size_t)b] + 1 - q.start[(size_t)a] < 1000000000) ++b;
				if (b == a) b = a + 1;
				const int64_t lo = a + (b - a) * k / ngpu, hi = a + (b - a) * (k + 1) / ngpu;      // this device's reads of the part
				a = b;
				if (hi <= lo) continue;
				std::vector<int32_t> qstart((size_t)(hi - lo));
				for (int64_t r = lo; r < hi; ++r) qstart[(size_t)(r - lo)] = (int32_t)(q.start[(size_t)r] - q.start[(size_t)lo]);
				mecat_asm_reads Q;
				Q.text = q.text.data() + q.start[(size_t)lo]; Q.num_letters = q.start[(size_t)(hi - 1)] + q.len[(size_t)(hi - 1)] + 1 - q.start[(size_t)lo];
				Q.num_reads = (int32_t)(hi - lo); Q.first_read_id = file_first[i - 1] + (int32_t)lo;
				Q.read_start = qstart.data(); Q.read_len = q.len.data() + lo;
				mecat_asm_overlap* ov = NULL;
				size_t n = 0;
				const double m0 = now();
				if (mecat_b200_asm_overlaps(ctx, idx, &Q, &P, &ov, &n)) { fprintf(stderr, "%s: %s\n", base, mecat_b200_last_error(ctx)); rcs[(size_t)k] = 1; break; }
				const double m1 = now();
				t_map[(size_t)k] += m1 - m0;
				{
					mbfmt::TextBuf tb;
					tb.s.reserve(n * 64);
					mbfmt::format_asm(tb, ov, n);
					std::lock_guard<std::mutex> g(out_mu[(size_t)slot]);
					if (fwrite(tb.s.data(), 1, tb.s.size(), out[(size_t)slot]) != tb.s.size()) rcs[(size_t)k] = 2;
			}
			totals[(size_t)k] += n;
			mecat_b200_free(ctx, ov);
			t_write[(size_t)k] += now() - m1;
			}
		}
		if (k == 0 && getenv("MECAT_B200_STATS")) {       // per-kernel CUDA-event times of device 0's share, one line
			mecat_b200_stats st;
			if (!mecat_b200_get_stats(ctx, &st))
				fprintf(stderr, "[kernel ms] asm_index=%.1f(%lld) asm_seed=%.1f(%lld) asm_extend=%.1f(%lld) hits=%lld candidates=%lld records=%lld h2d=%.1fMB d2h=%.1fMB\n",
				        st.kernel_ms[MECAT_K_ASM_INDEX], (long long)st.kernel_launches[MECAT_K_ASM_INDEX], st.kernel_ms[MECAT_K_ASM_SEED],
				        (long long)st.kernel_launches[MECAT_K_ASM_SEED], st.kernel_ms[MECAT_K_ASM_EXTEND], (long long)st.kernel_launches[MECAT_K_ASM_EXTEND],
				        (long long)st.num_hits, (long long)st.num_candidates, (long long)st.num_records, st.h2d_bytes / 1e6, st.d2h_bytes / 1e6);
		}
		if (!fast_exit) { mecat_b200_asm_index_release(ctx, idx); mecat_b200_destroy(ctx); }
	};
	if (ngpu == 1) work(0);
	else {
		std::vector<std::thread> th;
		for (int k = 0; k < ngpu; ++k) th.emplace_back(work, k);
		for (auto& t : th) t.join();
	}
	int rc = 0;
	size_t total = 0;
	double td = 0, ti = 0, tm = 0, tw = 0;
	for (int k = 0; k < ngpu; ++k) {
		if (rcs[(size_t)k]) rc = 1;
		if (rcs[(size_t)k] == 2) fprintf(stderr, "%s: writing the result failed\n", base);
		total += totals[(size_t)k];
		if (t_dev[(size_t)k] > td) td = t_dev[(size_t)k];
		if (t_index[(size_t)k] > ti) ti = t_index[(size_t)k];
		if (t_map[(size_t)k] > tm) tm = t_map[(size_t)k];
		if (t_write[(size_t)k] > tw) tw = t_write[(size_t)k];
	}
	for (FILE* f : out) if (fclose(f) != 0) { if (!rc) fprintf(stderr, "%s: writing the result failed\n", base); rc = 1; }
	if (!rc) fprintf(stderr, "[%s] load and driver start-up %.2f s, device context %.2f s, index %.2f s, mapping %.2f s, result files %.2f s (slowest of %d device%s), total %.2f s, %zu overlaps\n", base, t_init - t0,
	                 td, ti, tm, tw, ngpu, ngpu == 1 ? "" : "s", now() - t0, total);
	if (fast_exit) { fflush(stdout); fflush(stderr); _exit(rc); }
	return rc;
}
