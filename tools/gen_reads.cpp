// tools/gen_reads.cpp -- TEST/BENCH TOOLING (synthetic PacBio-CLR-like reads).
//
//   gen_reads <out.fasta> <num_reads> <genome_len> <seed> [mean_len=15000] [sd_len=1500]
//             [err=0.15] [genome_out.fasta | -] [first_read=0]
//
// first_read: ordinal of the first read written (reads first_read .. first_read+num_reads-1 of the
// same infinite, seed-defined read sequence), so several processes can each write their own slice.
//
// genome  : iid uniform ACGT of length G, base i = splitmix64(seed, i) & 3 (counter based).
// read r  : own xoshiro256** stream seeded by (seed, r) so the output does not depend on
//           the number of worker threads; length max(2000, round(N(mean, sd^2))) capped to G;
//           uniform start; per template base: deletion w.p. 0.30*err, substitution (to one of
//           the 3 other bases) w.p. 0.10*err, then geometric insertions w.p. 0.60*err each;
//           the read is reverse-complemented w.p. 0.5.  Pure ACGT, header = 0-based ordinal.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static inline uint64_t splitmix64(uint64_t x)
{
	x += 0x9E3779B97F4A7C15ULL;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
	return x ^ (x >> 31);
}

struct Rng
{
	uint64_t s[4];
	Rng(uint64_t seed, uint64_t stream)
	{
		uint64_t z = splitmix64(seed * 0x2545F4914F6CDD1DULL + stream);
		for (int i = 0; i < 4; ++i) { z = splitmix64(z); s[i] = z; }
	}
	static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
	uint64_t next()
	{
		uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
		s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
		return r;
	}
	double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

int main(int argc, char** argv)
{
	if (argc < 5) {
		fprintf(stderr, "usage: %s out.fasta num_reads genome_len seed [mean sd err genome.fasta]\n", argv[0]);
		return 1;
	}
	const char* out_path = argv[1];
	const long n_reads = atol(argv[2]);
	const long G = atol(argv[3]);
	const uint64_t seed = strtoull(argv[4], 0, 10);
	const double mean = argc > 5 ? atof(argv[5]) : 15000.0;
	const double sd = argc > 6 ? atof(argv[6]) : 1500.0;
	const double err = argc > 7 ? atof(argv[7]) : 0.15;
	const char* genome_out = (argc > 8 && strcmp(argv[8], "-") != 0) ? argv[8] : NULL;
	const long first_read = argc > 9 ? atol(argv[9]) : 0;
	const double p_del = 0.30 * err, p_sub = 0.10 * err, p_ins = 0.60 * err;

	std::vector<uint8_t> genome(G);
	int nt = (int)std::thread::hardware_concurrency();
	if (nt < 1) nt = 1;
	if (nt > 64) nt = 64;
	{
		std::vector<std::thread> th;
		for (int t = 0; t < nt; ++t)
			th.emplace_back([&, t]() {
				long lo = G * t / nt, hi = G * (t + 1) / nt;
				for (long i = lo; i < hi; ++i) genome[i] = splitmix64(seed ^ (0xA5A5ULL << 48) ^ (uint64_t)i) & 3;
			});
		for (auto& x : th) x.join();
	}
	if (genome_out) {
		FILE* g = fopen(genome_out, "w");
		if (!g) { perror(genome_out); return 1; }
		fprintf(g, ">genome\n");
		std::string line(G, 'A');
		for (long i = 0; i < G; ++i) line[i] = "ACGT"[genome[i]];
		fwrite(line.data(), 1, G, g);
		fputc('\n', g);
		fclose(g);
	}

	FILE* out = fopen(out_path, "w");
	if (!out) { perror(out_path); return 1; }
	const long CH = 4096;  // reads per round, generated in parallel then written in order
	std::vector<std::string> buf(CH);
	for (long base = 0; base < n_reads; base += CH) {
		long cnt = std::min(CH, n_reads - base);
		std::vector<std::thread> th;
		for (int t = 0; t < nt; ++t)
			th.emplace_back([&, t]() {
				for (long j = t; j < cnt; j += nt) {
					long r = first_read + base + j;
					Rng rng(seed, (uint64_t)r + 1);
					double u1 = rng.uni(), u2 = rng.uni();
					if (u1 < 1e-300) u1 = 1e-300;
					double z = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
					long len = lround(mean + sd * z);
					if (len < 2000) len = 2000;
					if (len > G) len = G;
					long start = (long)(rng.uni() * (double)(G - len + 1));
					if (start > G - len) start = G - len;
					bool flip = rng.uni() < 0.5;
					std::string s;
					s.reserve((size_t)(len * 1.15) + 64);
					for (long i = 0; i < len; ++i) {
						int b = genome[start + i];
						double u = rng.uni();
						if (u < p_del) {
						} else if (u < p_del + p_sub) {
							s.push_back((char)((b + 1 + (int)(rng.next() % 3)) & 3));
						} else
							s.push_back((char)b);
						while (rng.uni() < p_ins) s.push_back((char)(rng.next() & 3));
					}
					size_t L = s.size();
					std::string& o = buf[j];
					char hdr[32];
					int hl = snprintf(hdr, sizeof hdr, ">%ld\n", r);
					o.assign(hdr, hl);
					o.resize(hl + L + 1);
					char* d = &o[hl];
					if (!flip) for (size_t i = 0; i < L; ++i) d[i] = "ACGT"[(int)s[i]];
					else for (size_t i = 0; i < L; ++i) d[i] = "TGCA"[(int)s[L - 1 - i]];
					d[L] = '\n';
				}
			});
		for (auto& x : th) x.join();
		for (long j = 0; j < cnt; ++j) fwrite(buf[j].data(), 1, buf[j].size(), out);
	}
	fclose(out);
	return 0;
}
