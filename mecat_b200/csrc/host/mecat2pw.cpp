// mecat_b200/csrc/host/mecat2pw.cpp -- host driver with the reference's mecat2pw command line.
//
// Same flags, same work-directory protocol and same output text as the reference
//   main / merge_results           src/mecat2pw/pw.cpp:34-85
//   parse_arguments                src/mecat2pw/pw_options.cpp:73-211
//   process_one_volume             src/mecat2pw/pw_impl.cpp:835-882
//   operator<<(ExtensionCandidate) src/common/alignment.cpp:18-32
//   output_m4record                src/mecat2pw/pw_impl.cpp:509-531
// but every hot loop runs on the GPU through the C ABI (include/mecat_b200.h).  `-t` is
// accepted and ignored (there are no CPU worker threads).  GPUs: MECAT_GPUS=n (default 1): one
// host thread per device, the tiles (index volume s, query volume v >= s) are handed out one by
// one, so the devices stay balanced although row s has num_volumes - s tiles; the text of a row
// becomes the same wrk/r_N file the reference writes, so resume works unchanged.
#include <dirent.h>
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
#include <functional>
#include <future>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "mecat_b200.h"
#include "format.h"

namespace {

struct Options
{
	int task = 1;
	const char* reads = NULL;
	const char* output = NULL;
	const char* wrk_dir = NULL;
	int num_threads = 1;
	int num_candidates = 100;
	int min_align_size = 2000;
	int min_kmer_match = 4;
	int output_gapped_start_point = 0;
	int tech = 0;
};

void print_usage(const char* prog)
{
	fprintf(stderr, "\n\nusage:\n%s [-j task] [-d dataset] [-o output] [-w working dir] [-t threads] [-n candidates] [-g 0/1]\n\n", prog);
	fprintf(stderr, "options:\n");
	fprintf(stderr, "-j <integer>\tjob: 0 = seeding, 1 = align\n\t\tdefault: 1\n");
	fprintf(stderr, "-d <string>\treads file name\n");
	fprintf(stderr, "-o <string>\toutput file name\n");
	fprintf(stderr, "-w <string>\tworking folder name, will be created if not exist\n");
	fprintf(stderr, "-t <integer>\tnumber of cput threads (accepted, unused: the hot path runs on the GPU)\n\t\tdefault: 1\n");
	fprintf(stderr, "-n <integer>\tnumber of candidates for gapped extension\n\t\tDefault: 100\n");
	fprintf(stderr, "-a <integer>\tminimum size of overlaps\n\t\tDefault: 2000 if x = 0, 500 if x = 1\n");
	fprintf(stderr, "-k <integer>\tminimum number of kmer match a matched block has\n\t\tDefault: 4 if x = 0, 2 if x = 1\n");
	fprintf(stderr, "-g <0/1>\twhether print gapped extension start point, 0 = no, 1 = yes\n\t\tDefault: 0\n");
	fprintf(stderr, "-x <0/x>\tsequencing technology: 0 = pacbio, 1 = nanopore\n\t\tDefault: 0\n");
}

int parse_arguments(int argc, char* argv[], Options* o)
{
	int task = -1, num_threads = -1, num_candidates = -1, min_align_size = -1, min_kmer_match = -1, gapped = -1, tech = 0;
	int c;
	opterr = 0;
	while ((c = getopt(argc, argv, "j:d:o:w:t:n:g:x:a:k:")) != -1) {
		switch (c) {
		case 'j': task = atoi(optarg); break;
		case 'd': o->reads = optarg; break;
		case 'o': o->output = optarg; break;
		case 'w': o->wrk_dir = optarg; break;
		case 't': num_threads = atoi(optarg); break;
		case 'n': num_candidates = atoi(optarg); break;
		case 'a': min_align_size = atoi(optarg); break;
		case 'k': min_kmer_match = atoi(optarg); break;
		case 'g':
			if (optarg[0] == '0') gapped = 0;
			else if (optarg[0] == '1') gapped = 1;
			else { fprintf(stderr, "argument to option '-g' must be either '0' or '1'\n"); return 1; }
			break;
		case 'x':
			if (optarg[0] == '0') tech = 0;
			else if (optarg[0] == '1') tech = 1;
			else { fprintf(stderr, "invalid argument to option 'x': %s\n", optarg); abort(); }
			break;
		case '?': fprintf(stderr, "unrecognised option '%c'\n", (char)optopt); return 1;
		case ':': fprintf(stderr, "argument to option '%c' is not provided!\n", (char)optopt); return 1;
		}
	}
	o->tech = tech;
	if (tech == 1) { o->min_align_size = 500; o->min_kmer_match = 2; }
	if (task != -1) o->task = task;
	if (num_threads != -1) o->num_threads = num_threads;
	if (num_candidates != -1) o->num_candidates = num_candidates;
	if (min_align_size != -1) o->min_align_size = min_align_size;
	if (min_kmer_match != -1) o->min_kmer_match = min_kmer_match;
	if (gapped != -1) o->output_gapped_start_point = gapped;
	int ret = 0;
	if (o->task != 0 && o->task != 1) { fprintf(stderr, "task (-j) must be 0 or 1, not %d.\n", o->task); ret = 1; }
	if (!o->reads) { fprintf(stderr, "dataset must be specified.\n"); ret = 1; }
	else if (!o->output) { fprintf(stderr, "output must be specified.\n"); ret = 1; }
	else if (!o->wrk_dir) { fprintf(stderr, "working directory must be specified.\n"); ret = 1; }
	else if (o->num_threads < 1) { fprintf(stderr, "number of cpu threads must be > 0.\n"); ret = 1; }
	else if (o->num_candidates < 1) { fprintf(stderr, "number of candidates must be > 0.\n"); ret = 1; }
	if (ret) return ret;
	DIR* dir = opendir(o->wrk_dir);
	if (!dir) {
		if (mkdir(o->wrk_dir, S_IRWXU) == -1) { fprintf(stderr, "fail to create folder '%s'!\n", o->wrk_dir); exit(1); }
	} else closedir(dir);
	return 0;
}

struct StderrTimer   // DynamicTimer, src/common/defs.h:175-191
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

std::string results_name(const char* wrk, int vid, bool working)
{
	std::string n(wrk);
	if (n[n.size() - 1] != '/') n += '/';
	std::ostringstream os;
	os << "r_" << vid;
	if (working) os << ".working";
	return n + os.str();
}

int g_format_threads = 1;      // host threads a device thread may use to write its tile's text (set in main)

void format_records(std::string& text, const Options& opt, const void* rec, size_t n)
{
	auto piece = [&](size_t lo, size_t hi, std::string& out) {
		mbfmt::TextBuf b;
		b.s.reserve((hi - lo) * (opt.task == 0 ? 48 : 96) + 64);
		if (opt.task == 0) mbfmt::format_candidates(b, (const mecat_candidate*)rec + lo, hi - lo);
		else mbfmt::format_m4(b, (const mecat_m4*)rec + lo, hi - lo, opt.output_gapped_start_point != 0);
		out.swap(b.s);
	};
	const int T = n < 200000 ? 1 : g_format_threads;
	if (T <= 1) { piece(0, n, text); return; }
	// lines are independent: T slices formatted side by side, then joined in order
	std::vector<std::string> parts((size_t)T);
	std::vector<std::thread> pool;
	for (int t = 1; t < T; ++t) pool.emplace_back(piece, n * (size_t)t / (size_t)T, n * (size_t)(t + 1) / (size_t)T, std::ref(parts[(size_t)t]));
	piece(0, n / (size_t)T, parts[0]);
	for (auto& th : pool) th.join();
	size_t total = 0;
	for (const std::string& p : parts) total += p.size();
	text.clear();
	text.reserve(total);
	for (const std::string& p : parts) text.append(p);
}

// What one device keeps between tiles: the index volume it worked on last (create_ref_index of process_one_volume,
// pw_impl.cpp:855) -- consecutive tiles of the same row reuse it.
struct DeviceState
{
	mecat_b200_ctx* ctx = NULL;
	int svid = -1;
	mecat_volume ref;
	void* dref = NULL;
	void* index = NULL;
	void drop()
	{
		if (index) mecat_b200_index_release(ctx, index);
		if (dref) mecat_b200_volume_release(ctx, dref);
		if (svid >= 0) mecat_b200_volume_unload(&ref);
		index = NULL; dref = NULL; svid = -1;
	}
	bool use(int s, const std::vector<std::string>& vols)
	{
		if (svid == s) return true;
		drop();
		if (mecat_b200_volume_load(vols[(size_t)s].c_str(), &ref)) { fprintf(stderr, "failed to open file '%s'.\n", vols[(size_t)s].c_str()); return false; }
		svid = s;
		StderrTimer t("create_ref_index");
		if (mecat_b200_volume_upload(ctx, &ref, &dref) == 0 && mecat_b200_index_build(ctx, dref, &index) == 0) return true;
		fprintf(stderr, "mecat2pw: %s\n", mecat_b200_last_error(ctx));
		return false;
	}
};

// A query volume read from disk ahead of its tile (while the device is busy with the tile before it).
struct Loaded
{
	int vid = -1;
	bool ok = false;
	mecat_volume vol;
};

Loaded load_query_volume(int vid, const std::vector<std::string>& vols)
{
	Loaded l;
	l.vid = vid;
	memset(&l.vol, 0, sizeof l.vol);
	l.ok = mecat_b200_volume_load(vols[(size_t)vid].c_str(), &l.vol) == 0;
	return l;
}

// One tile of process_one_volume's loop (pw_impl.cpp:859-879): query volume vid against the index of volume svid.
// `pre`: the query volume if it was read ahead (consumed here), else it is read now.
bool process_tile(DeviceState& D, const Options& opt, int svid, int vid, const std::vector<std::string>& vols, Loaded* pre, std::string& text)
{
	Loaded q;
	if (pre && pre->vid == vid) { q = *pre; pre->vid = -1; pre->ok = false; }
	struct Unloader { Loaded& l; ~Unloader() { if (l.ok) mecat_b200_volume_unload(&l.vol); } } unloader{q};
	if (!D.use(svid, vols)) return false;
	mecat_pw_params p = {opt.task, opt.num_candidates, opt.min_align_size, opt.min_kmer_match, opt.tech};
	char info[64];
	snprintf(info, sizeof info, "process volume %d", vid);
	StderrTimer t(info);
	fprintf(stderr, "processing %s\n", vols[(size_t)vid].c_str());
	void* dreads = D.dref;
	bool ok = true;
	if (vid != svid) {
		if (q.vid != vid) q = load_query_volume(vid, vols);
		if (!q.ok) { fprintf(stderr, "failed to open file '%s'.\n", vols[(size_t)vid].c_str()); return false; }
		dreads = NULL;
		ok = mecat_b200_volume_upload(D.ctx, &q.vol, &dreads) == 0;
	}
	void* rec = NULL;
	size_t n = 0;
	// the lines of the tile are written on the device (MECAT_B200_TEXT=host: records to the host, printed by host threads)
	static const bool host_text = getenv("MECAT_B200_TEXT") && !strcmp(getenv("MECAT_B200_TEXT"), "host");
	if (ok && !host_text) {
		char* lines = NULL;
		size_t bytes = 0;
		ok = mecat_b200_pw_tile_text(D.ctx, D.index, D.dref, dreads, &p, opt.output_gapped_start_point != 0, &lines, &bytes, &n) == 0;
		if (ok) text.assign(lines ? lines : "", bytes);
		mecat_b200_free(D.ctx, lines);
	} else if (ok) {
		ok = mecat_b200_pw_tile(D.ctx, D.index, D.dref, dreads, &p, &rec, &n) == 0;
		if (ok) format_records(text, opt, rec, n);
	}
	if (!ok) fprintf(stderr, "mecat2pw: %s\n", mecat_b200_last_error(D.ctx));
	mecat_b200_free(D.ctx, rec);
	if (vid != svid && dreads) mecat_b200_volume_release(D.ctx, dreads);
	return ok;
}

// The tiles (s, v >= s) of one index volume and their text; r_s is written when the last one is in.
struct Row
{
	int svid = 0;
	std::vector<std::string> text;
	std::atomic<int> left{0};
};

}  // namespace

int main(int argc, char* argv[])
{
	Options opt;
	if (parse_arguments(argc, argv, &opt)) { print_usage(argv[0]); return 1; }
	// Creating a CUDA context takes one to two seconds; all devices' contexts come up next to the FASTA split.
	int want_gpus = 1;
	if (const char* g = getenv("MECAT_GPUS")) want_gpus = std::max(1, atoi(g));
	const int have = mecat_b200_device_count();
	if (have < 1) { fprintf(stderr, "mecat2pw: no CUDA device found (this build has no CPU path)\n"); return 1; }
	want_gpus = std::min(want_gpus, have);
	std::vector<mecat_b200_ctx*> warm_ctx((size_t)want_gpus, (mecat_b200_ctx*)NULL);
	std::vector<std::thread> warm;
	for (int d = 0; d < want_gpus; ++d) warm.emplace_back([&, d]() { if (mecat_b200_init(&warm_ctx[(size_t)d], d, NULL)) warm_ctx[(size_t)d] = NULL; });
	struct Joiner { std::vector<std::thread>& t; ~Joiner() { for (auto& x : t) if (x.joinable()) x.join(); } } warm_joiner{warm};
	int num_vols = 0;
	{
		StderrTimer t("split_raw_dataset");
		char err[512];
		const char* cap = getenv("MECAT_VOLUME_BASES");   // test hook: smaller volumes (reference: MCS, split_database.h:6-7)
		if (mecat_b200_split_dataset(opt.reads, opt.wrk_dir, cap ? atoll(cap) : 0, &num_vols, err, sizeof err)) {
			fprintf(stderr, "%s\n", err);
			return 1;
		}
	}
	std::vector<std::string> vols;
	{
		std::string idx(opt.wrk_dir);
		if (idx[idx.size() - 1] != '/') idx += '/';
		idx += "fileindex.txt";
		std::cout << idx << "\n";
		std::ifstream in(idx.c_str());
		std::string l;
		while (std::getline(in, l)) { if (!l.empty() && l[l.size() - 1] == '\r') l.erase(l.size() - 1); if (!l.empty()) vols.push_back(l); }
	}
	if ((int)vols.size() != num_vols) { fprintf(stderr, "volume index is inconsistent\n"); return 1; }

	int ngpus = want_gpus;
	{
		const int hw = (int)std::thread::hardware_concurrency();
		g_format_threads = std::max(1, std::min(8, (hw > 0 ? hw : 1) / std::max(1, ngpus)));
	}
	// Work items are tiles, row by row; rows whose r_N exists are finished (the reference's resume protocol, pw.cpp:65-81).
	// Any device takes the next tile: building the index of a volume costs a fraction of a tile, so several devices
	// share a row instead of each owning rows of very different sizes (row s has num_vols - s tiles).
	std::vector<Row> rows((size_t)num_vols);
	std::vector<std::pair<int, int>> tiles;
	for (int i = 0; i < num_vols; ++i) {
		rows[(size_t)i].svid = i;
		if (access(results_name(opt.wrk_dir, i, false).c_str(), F_OK) == 0) { fprintf(stderr, "volume %d has been finished\n", i); continue; }
		rows[(size_t)i].text.resize((size_t)(num_vols - i));
		rows[(size_t)i].left = num_vols - i;
		for (int v = i; v < num_vols; ++v) tiles.push_back(std::make_pair(i, v));
	}
	if (ngpus > (int)tiles.size()) ngpus = tiles.empty() ? 1 : (int)tiles.size();

	for (auto& x : warm) if (x.joinable()) x.join();
	for (int d = ngpus; d < want_gpus; ++d) if (warm_ctx[(size_t)d]) { mecat_b200_destroy(warm_ctx[(size_t)d]); warm_ctx[(size_t)d] = NULL; }
	std::atomic<int> next(0), failed(0);
	auto worker = [&](int dev) {
		DeviceState D;
		Loaded held;                                   // the volume read ahead for the tile this device takes next
		D.ctx = warm_ctx[(size_t)dev];
		if (!D.ctx) { fprintf(stderr, "mecat2pw: cannot initialise GPU %d\n", dev); failed = 1; return; }
		// A device holds one tile ahead of the one it works on, so that the query volume of the next tile is read from
		// disk while the device is busy.
		std::future<Loaded> ahead;
		int i = next.fetch_add(1);
		while (i < (int)tiles.size() && !failed) {
			const int j = next.fetch_add(1);
			if (j < (int)tiles.size() && tiles[(size_t)j].first != tiles[(size_t)j].second)
				ahead = std::async(std::launch::async, load_query_volume, tiles[(size_t)j].second, std::cref(vols));
			const int s = tiles[(size_t)i].first, v = tiles[(size_t)i].second;
			Row& row = rows[(size_t)s];
			Loaded pre = held;
			held = Loaded();
			if (!process_tile(D, opt, s, v, vols, &pre, row.text[(size_t)(v - s)])) { failed = 1; break; }
			if (row.left.fetch_sub(1) == 1) {          // the row is complete: r_s.working, then r_s
				const std::string working = results_name(opt.wrk_dir, s, true), done = results_name(opt.wrk_dir, s, false);
				std::ofstream out(working.c_str(), std::ios::binary);
				if (!out) { fprintf(stderr, "cannot open '%s' for writing\n", working.c_str()); failed = 1; break; }
				for (std::string& t : row.text) { out.write(t.data(), (std::streamsize)t.size()); std::string().swap(t); }
				out.close();
				if (!out || rename(working.c_str(), done.c_str()) != 0) { failed = 1; break; }
			}
			if (ahead.valid()) held = ahead.get();
			i = j;
		}
		if (ahead.valid()) held = ahead.get();
		if (held.ok) mecat_b200_volume_unload(&held.vol);
		D.drop();
	};
	std::vector<std::thread> th;
	for (int d = 1; d < ngpus; ++d) th.emplace_back(worker, d);
	worker(0);
	for (auto& t : th) t.join();
	if (failed) return 1;

	// merge_results: r_0 .. r_{n-1} concatenated in volume order (pw.cpp:34-46)
	std::unique_ptr<StderrTimer> merge_timer(new StderrTimer("merge_results"));
	std::ofstream merged(opt.output, std::ios::binary);
	if (!merged) { fprintf(stderr, "cannot open '%s' for writing\n", opt.output); return 1; }
	for (int i = 0; i < num_vols; ++i) {
		std::ifstream in(results_name(opt.wrk_dir, i, false).c_str(), std::ios::binary);
		if (in.peek() != std::ifstream::traits_type::eof()) merged << in.rdbuf();
	}
	merged.close();
	merge_timer.reset();
	if (!merged) { fprintf(stderr, "cannot write '%s'\n", opt.output); return 1; }
	// The contexts go last, device 0 last of all.  Returning tens of GB of pooled device memory block by block takes
	// ~1.5 s per device; everything this process owns is written and closed by now, so with MECAT_B200_FAST_EXIT unset or
	// 1 the process ends here and the driver reclaims the devices at once (MECAT_B200_FAST_EXIT=0: explicit release).
	const char* fe = getenv("MECAT_B200_FAST_EXIT");
	if (!fe || atoi(fe) != 0) {
		fflush(stdout); fflush(stderr);
		_exit(0);
	}
	{
		StderrTimer t("gpu release");
		for (int d = ngpus - 1; d >= 0; --d) if (warm_ctx[(size_t)d]) mecat_b200_destroy(warm_ctx[(size_t)d]);
	}
	return 0;
}
