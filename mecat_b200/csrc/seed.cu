// mecat_b200/csrc/seed.cu -- block k-mer seeding, DDF scoring and candidate selection (A2-A6).
//
// Replaces, for a whole query volume at once, the reference's per-read
//   extract_kmers / seeding / insert_loc      src/mecat2pw/pw_impl.cpp:83-97, 241-286, 121-159
//   find_location / get_candidates            src/mecat2pw/pw_impl.cpp:161-239, 288-465
//
// The reference scatters every index hit of a read strand into a dense array of 168-byte
// buckets (one per 2000 bases of the index volume) and then walks the touched buckets.  At
// 1.5 Gbase that is ~35 k random bucket updates per strand of which < 0.1 % ever matter.
// Here each strand is handled in three kernels:
//
//   k_seed   one CTA per (read, strand).  Streams the strand's hit lists three times:
//            (1) hashed 16-bit counters in shared memory count hits per bucket,
//            (2) buckets whose counter pair could pass the reference's `index_score >= 2k`
//                gate mark themselves and the +-ceil(L/2000) buckets a candidate's neighbour
//                vote can reach in a hashed interest bitmap,
//            (3) only hits of interesting buckets are collected, sorted by (bucket, k-mer
//                ordinal, offset) and turned into exact bucket records -- including the
//                order-dependent insert_loc eviction for buckets that overflow 40 seeds and
//                the time-of-update `index_score` snapshot -- written to an HBM arena.
//            Hash collisions only add buckets (a superset); everything downstream is exact.
//   k_walk   one warp per (read, strand): the sequential get_candidates walk in first-touch
//            order over the bucket records, with find_location / neighbour votes lane parallel.
//   k_merge  one warp per read: stable merge of the forward and reverse candidate lists.
#include "common.cuh"

namespace mb {

namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int CNT_SLOTS = 65536;            // hashed hit counters, slot = bucket & 0xFFFF
constexpr int BIT_WORDS = 2048;             // hashed interest bitmap (65536 bits) / anchor claim slots
constexpr uint32_t CNT_SAT = 0x8000u;       // 16-bit counters stop growing here (no wrap with <= 1024 racing adds)
constexpr uint32_t CNT_SAT8 = 128u;         // 8-bit counters: an add that finds 128 or more takes itself back
// Two shapes of k_seed (template parameter CTAS = CTAs per SM):
//   1: one CTA of 1 024 threads per SM, 16-bit counters (128 KB), 8 192 collected hits sort in shared memory;
//   2: two CTAs of 512 threads per SM, 8-bit counters (64 KB), 4 096 hits sort in shared memory.  The kernel is a chain of
//      block-wide phases separated by barriers (ncu at one CTA per SM: 3 of 11 stall cycles per issue are barrier waits,
//      issue slots 55 % busy); with two strands in flight per SM one CTA's barrier hides behind the other's gathers.
template <int CTAS> struct SeedShape
{
	static constexpr int THREADS = 1024 / CTAS;
	static constexpr int CNT_WORDS = CTAS == 1 ? CNT_SLOTS / 2 : CNT_SLOTS / 4;
	static constexpr int SCAP = 8192 / CTAS;   // collected hits that sort in shared memory (the counters' memory)
};
constexpr int KCACHE = 2048;               // sampled k-mers whose list info is cached in shared memory
constexpr int MAX_KM = 32766;               // seed ordinals are `short` in the reference (pw_impl.h:31)

struct BucketHdr          // 16 bytes, one per collected bucket, ascending seg inside a strand
{
	int32_t seg;
	int16_t score;        // Back_List::score (may exceed 40)
	int16_t iscore;       // index_score snapshot (pw_impl.cpp:270-280)
	uint16_t first_seed;  // seed ordinal (km+1) of the first accepted hit: first-touch order key
	uint16_t nst;         // stored seeds = min(accepted, 40)
	uint32_t eoff;        // first entry, relative to the strand's entry block
};

struct StrandDesc         // one per (read, strand) of a batch
{
	unsigned long long hdr_off;    // arena offsets in bytes
	unsigned long long ent_off;
	unsigned long long ord_off;
	int32_t nb;           // buckets
	int32_t nord;         // buckets with iscore >= gate, in first-touch order
	int32_t status;       // 0 ok, else error code
	int32_t pad;
};

struct SeedScratch        // per-CTA global scratch
{
	uint32_t* kb;         // list begin per sampled k-mer
	uint8_t* kc;          // list length per sampled k-mer
	unsigned long long* keys;   // collected hits
	unsigned long long* ent;    // deduplicated hits
	uint32_t* bstart;     // first entry of each bucket (+1 sentinel)
	unsigned long long* okeys;  // walk-order sort keys
};

// 13-mer code (first base most significant) of a query strand at sampled ordinal km
__device__ __forceinline__ uint32_t query_code(const uint32_t* __restrict__ arr, uint32_t g0, uint32_t comp, int km)
{
	return rev_groups2(ld_bases32(arr, g0 + (uint32_t)(km * STRIDE)) ^ comp) >> 6;
}

// block-wide bitonic sort of n (power of two) 64-bit keys.  Two strides are folded into one step
// (a thread holds the 4 keys i, i+h, i+2h, i+3h), which halves the barriers and the shared-memory
// round trips; all index arithmetic is shifts (strides are powers of two).
__device__ __forceinline__ void cmpex(unsigned long long& x, unsigned long long& y, bool up)
{
	if ((x > y) == up) { const unsigned long long t = x; x = y; y = t; }
}

__device__ void block_sort(unsigned long long* a, int n)
{
	for (int size = 2; size <= n; size <<= 1) {
		int ls = 31 - __clz(size) - 1;                 // log2 of the first stride, size / 2
		for (; ls >= 1; ls -= 2) {                     // strides 2^ls and 2^(ls-1) together
			const int lh = ls - 1, h = 1 << lh;
			__syncthreads();
			for (int t = threadIdx.x; t < (n >> 2); t += blockDim.x) {
				const int i0 = ((t >> lh) << (lh + 2)) | (t & (h - 1));
				const bool up = (i0 & size) == 0;
				unsigned long long k0 = a[i0], k1 = a[i0 + h], k2 = a[i0 + 2 * h], k3 = a[i0 + 3 * h];
				cmpex(k0, k2, up); cmpex(k1, k3, up);
				cmpex(k0, k1, up); cmpex(k2, k3, up);
				a[i0] = k0; a[i0 + h] = k1; a[i0 + 2 * h] = k2; a[i0 + 3 * h] = k3;
			}
		}
		if (ls == 0) {                                 // left-over stride 1
			__syncthreads();
			for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
				const int i = t << 1;
				const bool up = (i & size) == 0;
				unsigned long long x = a[i], y = a[i + 1];
				cmpex(x, y, up);
				a[i] = x; a[i + 1] = y;
			}
		}
	}
	__syncthreads();
}

// block-wide exclusive scan helper: returns the exclusive prefix of `v` over the block and
// the block total through *total (all threads must call)
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* wsum /* >= 33 ints smem */)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	int inc = v;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		int n = __shfl_up_sync(FULL, inc, off);
		if (lane >= off) inc += n;
	}
	__syncthreads();
	if (lane == 31) wsum[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		int w = lane < nw ? wsum[lane] : 0;
		int winc = w;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			int n = __shfl_up_sync(FULL, winc, off);
			if (lane >= off) winc += n;
		}
		if (lane < nw) wsum[lane] = winc - w;
		if (lane == 31) wsum[32] = winc;
	}
	__syncthreads();
	*total = wsum[32];
	return wsum[warp] + inc - v;
}

struct SeedParams
{
	const uint32_t* begin;
	const int32_t* pos;
	const uint32_t* qfwd;
	const uint32_t* qrev;
	const int2* qoffsz;
	int qN;
	int read0, nreads;      // batch
	int gate;               // 2 * min_kmer_match
	unsigned char* arena;
	unsigned long long arena_bytes;
	unsigned long long* arena_cursor;
	unsigned int* work_counter;
	StrandDesc* desc;
	SeedScratch* scratch;   // one per CTA
	int kcap;               // capacity of kb / kc
	int hcap;               // capacity of keys / ent / bstart / okeys
	unsigned long long* hit_counter;
};

// DDF pair predicate of insert_loc (pw_impl.cpp:135): i before j in slot order
__device__ __forceinline__ bool pair_ok(int li, int si, int lj, int sj)
{
	return (sj - si > 0) && (lj - li > 0) && ddf_close(lj - li, sj - si);
}

// One warp replays the acceptance sequence of one overflowing bucket (> 40 accepted hits):
// exact insert_loc semantics with incrementally maintained consistency counts.
// ent[0..na) are the accepted hits in time order; the score after each accept is written to
// the top 16 bits of the entry.  Final slots go to sl/ss (40 each); returns the final score.
__device__ int replay_overflow(unsigned long long* ent, int na, int* sl, int* ss, int* cc, int lane)
{
	// the first 40 accepts fill the slots in order
	for (int i = lane; i < SLOTS; i += 32) {
		const unsigned long long e = ent[i];
		sl[i] = (int)(e & 2047u);
		ss[i] = (int)((e >> 11) & 0xFFFFu) + 1;
		ent[i] = (e & 0x0000FFFFFFFFFFFFull) | ((unsigned long long)(i + 1) << 48);
	}
	__syncwarp();
	for (int x = lane; x < SLOTS; x += 32) {
		int c = 0;
		for (int y = 0; y < SLOTS; ++y)
			if (y != x) c += (y < x) ? pair_ok(sl[y], ss[y], sl[x], ss[x]) : pair_ok(sl[x], ss[x], sl[y], ss[y]);
		cc[x] = c;
	}
	__syncwarp();
	int score = SLOTS;
	for (int t = SLOTS; t < na; ++t) {
		// Fast forward.  While all 40 slots are mutually consistent (every count is 39), a newcomer
		// either agrees with all of them and overwrites slot 39, or disagrees with at least two and
		// is dropped; both leave slots 0..38 and all counts as they are and keep the score increment.
		// Only slot 39 (the last accepted hit) carries over, so 32 newcomers are judged at once; the
		// first one that would evict a slot (exactly one disagreement) goes through the general step.
		const bool clean = __all_sync(FULL, cc[lane] == SLOTS - 1 && (lane + 32 >= SLOTS || cc[lane + 32] == SLOTS - 1));
		if (clean) {
			int last_l = sl[SLOTS - 1], last_s = ss[SLOTS - 1];
			const int gap = score - t;       // score minus hits seen: every settled hit keeps its increment
			bool stop = false;
			while (t < na && !stop) {
				const int me = t + lane;
				const bool valid = me < na;
				const unsigned long long e = valid ? ent[me] : 0ull;
				const int nl = (int)(e & 2047u), ns = (int)((e >> 11) & 0xFFFFu) + 1;
				int v38 = 0;
				for (int x = 0; x < SLOTS - 1; ++x) v38 += (int)pair_ok(sl[x], ss[x], nl, ns);
				const unsigned amask = __ballot_sync(FULL, valid && v38 == SLOTS - 1);      // would-be accepts
				const unsigned below = amask & ((1u << lane) - 1u);
				const int src = below ? 31 - __clz(below) : 0;
				int pl = __shfl_sync(FULL, nl, src), ps = __shfl_sync(FULL, ns, src);
				if (!below) { pl = last_l; ps = last_s; }
				const bool plast = pair_ok(pl, ps, nl, ns);
				const bool viol = valid && ((v38 == SLOTS - 1 && !plast) || (v38 == SLOTS - 2 && plast));
				const unsigned vmask = __ballot_sync(FULL, viol);
				const int upto = vmask ? __ffs(vmask) - 1 : min(32, na - t);                 // hits settled here
				if (lane < upto) ent[me] = (e & 0x0000FFFFFFFFFFFFull) | ((unsigned long long)((me + 1 + gap) & 0xFFFF) << 48);
				const unsigned done = amask & (upto >= 32 ? FULL : ((1u << upto) - 1u));
				if (done) {
					const int w = 31 - __clz(done);
					last_l = __shfl_sync(FULL, nl, w); last_s = __shfl_sync(FULL, ns, w);
				}
				t += upto;
				stop = vmask != 0u;
			}
			__syncwarp();
			if (lane == 0) { sl[SLOTS - 1] = last_l; ss[SLOTS - 1] = last_s; }
			__syncwarp();
			score = t + gap;
			if (t >= na) break;
		}
		const unsigned long long e = ent[t];
		const int nl = (int)(e & 2047u), ns = (int)((e >> 11) & 0xFFFFu) + 1;
		++score;
		// votes: existing slot x gets cc[x] + P(x,new); the new entry gets sum P(x,new)
		const int x0 = lane, x1 = lane + 32;
		const bool p0 = pair_ok(sl[x0], ss[x0], nl, ns);
		const bool p1 = (x1 < SLOTS) ? pair_ok(sl[x1], ss[x1], nl, ns) : false;
		const int vnew = __popc(__ballot_sync(FULL, p0)) + __popc(__ballot_sync(FULL, p1));
		int key = ((cc[x0] + (int)p0) << 8) | x0;
		if (x1 < SLOTS) key = min(key, ((cc[x1] + (int)p1) << 8) | x1);
		key = min(key, (vnew << 8) | SLOTS);
		key = __reduce_min_sync(FULL, key);
		const int minval = key >> 8, mini = key & 255;
		if (minval == SLOTS) {
			// everything mutually consistent: the newcomer replaces slot 39, score keeps the increment
			const int ol = sl[SLOTS - 1], os = ss[SLOTS - 1];
			__syncwarp();
			int add0 = 0, add1 = 0;
			if (x0 < SLOTS - 1) add0 = (int)p0 - (int)pair_ok(sl[x0], ss[x0], ol, os);
			if (x1 < SLOTS - 1) add1 = (int)p1 - (int)pair_ok(sl[x1], ss[x1], ol, os);
			const bool p39 = __shfl_sync(FULL, (int)p1, SLOTS - 1 - 32) != 0;   // P(old39, new), lane 7 holds x1 = 39
			__syncwarp();
			if (x0 < SLOTS - 1) cc[x0] += add0;
			if (x1 < SLOTS - 1) cc[x1] += add1;
			if (lane == 0) { sl[SLOTS - 1] = nl; ss[SLOTS - 1] = ns; cc[SLOTS - 1] = vnew - (int)p39; }
		} else if (mini < SLOTS) {
			// evict slot `mini`, shift the rest down, newcomer becomes slot 39
			const int ml = sl[mini], ms = ss[mini];
			int l0 = sl[x0], s0 = ss[x0], c0 = cc[x0];
			int l1 = 0, s1 = 0, c1 = 0;
			if (x1 < SLOTS) { l1 = sl[x1]; s1 = ss[x1]; c1 = cc[x1]; }
			if (x0 != mini) c0 += (int)p0 - (int)((x0 < mini) ? pair_ok(l0, s0, ml, ms) : pair_ok(ml, ms, l0, s0));
			if (x1 < SLOTS && x1 != mini) c1 += (int)p1 - (int)((x1 < mini) ? pair_ok(l1, s1, ml, ms) : pair_ok(ml, ms, l1, s1));
			const unsigned b0 = __ballot_sync(FULL, p0), b1 = __ballot_sync(FULL, p1);
			const bool pm = mini < 32 ? ((b0 >> mini) & 1u) : ((b1 >> (mini - 32)) & 1u);
			__syncwarp();
			if (x0 != mini) { const int d = x0 > mini ? x0 - 1 : x0; sl[d] = l0; ss[d] = s0; cc[d] = c0; }
			if (x1 < SLOTS && x1 != mini) { const int d = x1 > mini ? x1 - 1 : x1; sl[d] = l1; ss[d] = s1; cc[d] = c1; }
			if (lane == 0) { sl[SLOTS - 1] = nl; ss[SLOTS - 1] = ns; cc[SLOTS - 1] = vnew - (int)pm; }
			--score;
		}
		// else: the newcomer itself is the least consistent -> dropped, score keeps the increment
		__syncwarp();
		if (lane == 0) ent[t] = (e & 0x0000FFFFFFFFFFFFull) | ((unsigned long long)(score & 0xFFFF) << 48);
	}
	__syncwarp();
	return score;
}

template <int CTAS>
__global__ void __launch_bounds__(SeedShape<CTAS>::THREADS, CTAS) k_seed(SeedParams P)
{
	constexpr int SCAP = SeedShape<CTAS>::SCAP;
	constexpr int CNT_WORDS = SeedShape<CTAS>::CNT_WORDS;
	extern __shared__ uint32_t smem_u32[];
	uint32_t* cnt = smem_u32;                         // CNT_WORDS words of 16-bit (CTAS = 1) or 8-bit hit counters
	uint32_t* want_bits = cnt + CNT_WORDS;            // BIT_WORDS words = 65536 interest bits, same slot map
	uint32_t* kbs = want_bits + BIT_WORDS;            // KCACHE list begins
	uint8_t* kcs = (uint8_t*)(kbs + KCACHE);          // KCACHE list lengths
	int* misc = (int*)(kcs + KCACHE);                 // 64 ints: scan scratch [0..33), counters
	int* wslots = misc + 64;                          // per warp: 3 x 41 ints for the overflow replay
	// once the counters are dead (after pass 2) their 128 KB hold the sort buffers
	unsigned long long* skeys = (unsigned long long*)cnt;          // SCAP collected hits
	unsigned long long* sent = skeys + SCAP;                       // SCAP accepted hits
	__shared__ unsigned int s_item;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
	const SeedScratch gsc = P.scratch[blockIdx.x];

	for (;;) {
		__syncthreads();
		if (tid == 0) s_item = atomicAdd(P.work_counter, 1u);
		__syncthreads();
		const unsigned item = s_item;
		if (item >= 2u * (unsigned)P.nreads) break;
		const int r = P.read0 + (int)(item >> 1), strand = (int)(item & 1u);
		const int2 qo = P.qoffsz[r];
		const int L = qo.y;
		StrandDesc D;
		D.hdr_off = D.ent_off = D.ord_off = 0; D.nb = 0; D.nord = 0; D.status = 0; D.pad = 0;
		const int nk = L >= KMER ? (L - KMER) / STRIDE + 1 : 0;
		if (nk > MAX_KM || nk > P.kcap) { if (tid == 0) { D.status = 2; P.desc[item] = D; } continue; }
		const uint32_t* arr = strand ? P.qrev : P.qfwd;
		const uint32_t g0 = strand ? (uint32_t)(P.qN - qo.x - L) : (uint32_t)qo.x;
		const uint32_t comp = strand ? FULL : 0u;
		const int W = (L + SEGW - 1) / SEGW;          // reach of a candidate's neighbour vote, in buckets
		// list begin / length per sampled k-mer: shared memory for ordinary reads, global scratch for very long ones
		uint32_t* kb = nk <= KCACHE ? kbs : gsc.kb;
		uint8_t* kc = nk <= KCACHE ? kcs : gsc.kc;

		// ---- clear
		for (int i = tid; i < CNT_WORDS + BIT_WORDS; i += blockDim.x) cnt[i] = 0u;
		if (tid < 64) misc[tid] = 0;
		// ---- k-mer lookup
		for (int km = tid; km < nk; km += blockDim.x) {
			const uint32_t code = query_code(arr, g0, comp, km);
			const uint32_t b = P.begin[code], e = P.begin[code + 1];
			kb[km] = b;
			kc[km] = (uint8_t)(e - b);
		}
		__syncthreads();
		// ---- pass 1: hit counts per bucket slot.  The slot map keeps neighbouring buckets adjacent,
		// aliases (buckets 65536 apart) only inflate counts: every later decision is a superset.
		// Four lists per warp iteration: their position loads are issued together (the passes are bound by
		// L2/DRAM latency of these gathers, not by bandwidth).
		unsigned myhits = 0;
		constexpr int U = 4;
		for (int km0 = warp; km0 < nk; km0 += U * nwarps) {
			int n[U];
			uint32_t b[U], p0[U], p1[U] = {0u, 0u, 0u, 0u};
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int km = km0 + u * nwarps;
				n[u] = km < nk ? (int)kc[km] : 0;
				b[u] = km < nk ? kb[km] : 0u;
			}
			// most lists hold <= 32 positions (mean ~26 at 1.5 Gbase): the second gather and everything behind it is skipped
			// for the whole warp unless one of the four lists is longer
			const bool long_list = max(max(n[0], n[1]), max(n[2], n[3])) > 32;
#pragma unroll
			for (int u = 0; u < U; ++u) p0[u] = lane < n[u] ? (uint32_t)P.pos[b[u] + lane] : 0u;
			if (long_list) {
#pragma unroll
				for (int u = 0; u < U; ++u) p1[u] = lane + 32 < n[u] ? (uint32_t)P.pos[b[u] + lane + 32] : 0u;
			}
			auto count_one = [&](uint32_t pp) {
				const uint32_t hh = (pp / SEGW) & (CNT_SLOTS - 1);
				if (CTAS == 1) {
					const uint32_t cur = (cnt[hh >> 1] >> ((hh & 1u) << 4)) & 0xFFFFu;
					if (cur < CNT_SAT) atomicAdd(&cnt[hh >> 1], 1u << ((hh & 1u) << 4));
				} else {
					// 8-bit counter: add, and take the add back when the counter had reached the cap.  Adds and their
					// take-backs are one integer addition on the word, so a transient carry into the neighbour byte
					// cancels exactly; the atomics of one counter return successive values, so it ends at <= 128.
					const uint32_t sh = (hh & 3u) << 3;
					if (((cnt[hh >> 2] >> sh) & 0xFFu) < CNT_SAT8) {
						const uint32_t old = atomicAdd(&cnt[hh >> 2], 1u << sh);
						if (((old >> sh) & 0xFFu) >= CNT_SAT8) atomicSub(&cnt[hh >> 2], 1u << sh);
					}
				}
				++myhits;
			};
#pragma unroll
			for (int u = 0; u < U; ++u) if (lane < n[u]) count_one(p0[u]);
			if (long_list) {
#pragma unroll
				for (int u = 0; u < U; ++u) {
					if (lane + 32 < n[u]) count_one(p1[u]);
					for (int h = lane + 64; h < n[u]; h += 32) count_one((uint32_t)P.pos[b[u] + h]);
				}
			}
		}
		__syncthreads();
		// ---- pass 2: scan the slot table.  A bucket can only pass the reference's index_score >= 2k
		// gate if its own count plus its left neighbour's reaches the gate; such a slot marks the
		// +-W slots a candidate anchored there can reach (previous bucket, neighbour votes).
		for (int w = tid; w < CNT_WORDS; w += blockDim.x) {
			const uint32_t cw = cnt[w];
			if (cw == 0u) continue;
			const uint32_t pw = cnt[(w + CNT_WORDS - 1) & (CNT_WORDS - 1)];
			constexpr int PER = CTAS == 1 ? 2 : 4;            // counters per word
			constexpr int BITS = 32 / PER;
			constexpr uint32_t CMASK = (1u << BITS) - 1u;
			int left = (int)(pw >> (32 - BITS));              // the last counter of the word before
#pragma unroll
			for (int q = 0; q < PER; ++q) {
				const int own = (int)((cw >> (q * BITS)) & CMASK);
				const int c = own + left;
				left = own;
				if (own == 0 || c < P.gate) continue;
				const int slot = PER * w + q;
				int lo = slot - W, hi = slot + W;              // inclusive, modulo CNT_SLOTS
				if (hi - lo + 1 >= CNT_SLOTS) { lo = 0; hi = CNT_SLOTS - 1; }
				for (int wb = (lo >> 5); wb <= (hi >> 5); ++wb) {
					const int first = max(lo, wb << 5) - (wb << 5), lastb = min(hi, (wb << 5) + 31) - (wb << 5);
					const uint32_t mask = (lastb == 31 ? 0xFFFFFFFFu : ((1u << (lastb + 1)) - 1u)) & ~((1u << first) - 1u);
					uint32_t* dst = &want_bits[wb & (BIT_WORDS - 1)];
					if ((*dst & mask) != mask) atomicOr(dst, mask);
				}
			}
		}
		__syncthreads();
		// ---- pass 3: collect the hits of wanted buckets (shared memory first, overflow to global scratch)
		for (int km0 = warp; km0 < nk; km0 += U * nwarps) {
			int n[U];
			uint32_t b[U], p0[U], p1[U] = {0u, 0u, 0u, 0u};
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int km = km0 + u * nwarps;
				n[u] = km < nk ? (int)kc[km] : 0;
				b[u] = km < nk ? kb[km] : 0u;
			}
			const bool long_list = max(max(n[0], n[1]), max(n[2], n[3])) > 32;
#pragma unroll
			for (int u = 0; u < U; ++u) p0[u] = lane < n[u] ? (uint32_t)P.pos[b[u] + lane] : 0u;
			if (long_list) {
#pragma unroll
				for (int u = 0; u < U; ++u) p1[u] = lane + 32 < n[u] ? (uint32_t)P.pos[b[u] + lane + 32] : 0u;
			}
			auto collect = [&](bool have, uint32_t p, int km) {
				bool take = false;
				if (have) {
					const uint32_t h2 = (p / SEGW) & (CNT_SLOTS - 1);
					take = (want_bits[h2 >> 5] >> (h2 & 31u)) & 1u;
				}
				const unsigned m = __ballot_sync(FULL, take);
				if (m) {
					int base = 0;
					if (lane == 0) base = atomicAdd(&misc[40], __popc(m));
					base = __shfl_sync(FULL, base, 0);
					if (take) {
						const int at = base + __popc(m & ((1u << lane) - 1u));
						const unsigned long long key = ((unsigned long long)(p / SEGW) << 27) | ((unsigned long long)km << 11) | (unsigned long long)(p % SEGW);
						if (at < SCAP) skeys[at] = key;          // the counters are dead: their memory takes the keys
						else if (at < P.hcap) gsc.keys[at] = key;
					}
				}
			};
#pragma unroll
			for (int u = 0; u < U; ++u) {
				if (n[u] == 0) continue;
				const int km = km0 + u * nwarps;
				collect(lane < n[u], p0[u], km);
				if (n[u] > 32) collect(lane + 32 < n[u], p1[u], km);
				for (int h0 = 64; h0 < n[u]; h0 += 32) {
					const int h = h0 + lane;
					collect(h < n[u], h < n[u] ? (uint32_t)P.pos[b[u] + h] : 0u, km);
				}
			}
		}
		if (P.hit_counter) {
			myhits = __reduce_add_sync(FULL, myhits);
			if (lane == 0 && myhits) atomicAdd(P.hit_counter, (unsigned long long)myhits);
		}
		__syncthreads();
		const int ncol = misc[40];
		if (ncol > P.hcap) { if (tid == 0) { D.status = 3; P.desc[item] = D; } continue; }
		if (ncol == 0) { if (tid == 0) P.desc[item] = D; continue; }
		// Ordinary strands sort and de-duplicate in the 128 KB the counters occupied; only strands with
		// more than SCAP collected hits fall back to the per-CTA global scratch.
		const bool in_smem = ncol <= SCAP;
		SeedScratch sc = gsc;
		if (in_smem) {
			sc.keys = skeys; sc.ent = sent;
			sc.bstart = (uint32_t*)skeys;                        // keys are dead once `ent` is built
		} else {
			for (int i = tid; i < SCAP; i += blockDim.x) gsc.keys[i] = skeys[i];
			__syncthreads();
		}
		int n2 = 1;
		while (n2 < ncol) n2 <<= 1;
		for (int i = ncol + tid; i < n2; i += blockDim.x) sc.keys[i] = ~0ull;
		block_sort(sc.keys, n2);

		// ---- de-duplicate: only the first (lowest) position of a (bucket, k-mer) pair is accepted
		// (pw_impl.cpp:265,282); compact accepted hits into `ent`
		int nent = 0;
		for (int base = 0; base < ncol; base += blockDim.x) {
			const int i = base + tid;
			int f = 0;
			unsigned long long k = 0;
			if (i < ncol) { k = sc.keys[i]; f = (i == 0) || ((sc.keys[i - 1] >> 11) != (k >> 11)); }
			int tot;
			const int ex = block_excl_scan(f, &tot, misc);
			if (f) sc.ent[nent + ex] = k;
			nent += tot;
		}
		__syncthreads();
		// ---- bucket boundaries
		int nb = 0;
		for (int base = 0; base < nent; base += blockDim.x) {
			const int i = base + tid;
			int f = 0;
			if (i < nent) f = (i == 0) || ((sc.ent[i - 1] >> 27) != (sc.ent[i] >> 27));
			int tot;
			const int ex = block_excl_scan(f, &tot, misc);
			if (f) sc.bstart[nb + ex] = (uint32_t)i;
			nb += tot;
		}
		if (tid == 0) sc.bstart[nb] = (uint32_t)nent;
		// walk-order keys: behind bstart in shared memory when few buckets, else global scratch
		if (in_smem && nb <= SCAP / 4) sc.okeys = skeys + SCAP / 2 + 1;
		__syncthreads();
		// ---- arena space: headers, entries (<= 40 per bucket), order list
		int nstore = 0;
		for (int base = 0; base < nb; base += blockDim.x) {
			const int b = base + tid;
			int v = 0;
			if (b < nb) v = min((int)(sc.bstart[b + 1] - sc.bstart[b]), SLOTS);
			int tot;
			block_excl_scan(v, &tot, misc);
			nstore += tot;
		}
		if (tid == 0) {
			const unsigned long long need = (unsigned long long)nb * sizeof(BucketHdr) + (unsigned long long)nstore * 4ull + (unsigned long long)nb * 4ull + 64ull;
			const unsigned long long at = atomicAdd(P.arena_cursor, (need + 15ull) & ~15ull);
			unsigned long long* a = (unsigned long long*)(misc + 44);
			a[0] = (at + need <= P.arena_bytes) ? at : ~0ull;
		}
		__syncthreads();
		const unsigned long long at = *(unsigned long long*)(misc + 44);
		if (at == ~0ull) { if (tid == 0) { D.status = 1; P.desc[item] = D; } continue; }
		D.hdr_off = at;
		D.ent_off = at + (unsigned long long)nb * sizeof(BucketHdr);
		D.ord_off = D.ent_off + (unsigned long long)nstore * 4ull;
		D.nb = nb;
		BucketHdr* hdr = (BucketHdr*)(P.arena + D.hdr_off);
		ushort2* ents = (ushort2*)(P.arena + D.ent_off);
		uint32_t* ord = (uint32_t*)(P.arena + D.ord_off);

		// ---- entry offsets per bucket (exclusive scan again, now storing)
		{
			int run = 0;
			for (int base = 0; base < nb; base += blockDim.x) {
				const int b = base + tid;
				int v = 0;
				if (b < nb) v = min((int)(sc.bstart[b + 1] - sc.bstart[b]), SLOTS);
				int tot;
				const int ex = block_excl_scan(v, &tot, misc);
				if (b < nb) hdr[b].eoff = (uint32_t)(run + ex);
				run += tot;
			}
		}
		__syncthreads();
		// ---- per bucket: final seeds and score (one warp per bucket)
		for (int b = warp; b < nb; b += nwarps) {
			const int s0 = (int)sc.bstart[b], na = (int)sc.bstart[b + 1] - s0;
			unsigned long long* e = sc.ent + s0;
			ushort2* out = ents + hdr[b].eoff;
			int score;
			if (na <= SLOTS) {
				for (int i = lane; i < na; i += 32) {
					const unsigned long long k = e[i];
					out[i] = make_ushort2((unsigned short)(k & 2047u), (unsigned short)(((k >> 11) & 0xFFFFu) + 1));
					e[i] = (k & 0x0000FFFFFFFFFFFFull) | ((unsigned long long)(i + 1) << 48);
				}
				score = na;
			} else {
				// More than 40 accepted hits.  Self hits (the read against its own copy in the index) are
				// exactly collinear: off = c + 10 * seed.  Then every pair is DDF consistent, every
				// insert_loc call sees minval == 40 and just overwrites slot 39, and the score never drops:
				// slots 0..38 = first 39 hits, slot 39 = last hit, score = number of accepted hits.
				bool col = true;
				const unsigned long long k00 = e[0];
				const int c0 = (int)(k00 & 2047u) - STRIDE * ((int)((k00 >> 11) & 0xFFFFu) + 1);
				for (int i = lane; i < na; i += 32) {
					const unsigned long long k = e[i];
					col = col && ((int)(k & 2047u) - STRIDE * ((int)((k >> 11) & 0xFFFFu) + 1) == c0);
				}
				if (__all_sync(FULL, col)) {
					for (int i = lane; i < na; i += 32) {
						const unsigned long long k = e[i];
						if (i < SLOTS - 1) out[i] = make_ushort2((unsigned short)(k & 2047u), (unsigned short)(((k >> 11) & 0xFFFFu) + 1));
						if (i == na - 1) out[SLOTS - 1] = make_ushort2((unsigned short)(k & 2047u), (unsigned short)(((k >> 11) & 0xFFFFu) + 1));
						e[i] = (k & 0x0000FFFFFFFFFFFFull) | ((unsigned long long)(i + 1) << 48);
					}
					score = na;
				} else {
					int* sl = wslots + warp * 123;
					score = replay_overflow(e, na, sl, sl + 41, sl + 82, lane);
					for (int i = lane; i < SLOTS; i += 32) out[i] = make_ushort2((unsigned short)sl[i], (unsigned short)sl[41 + i]);
				}
			}
			if (lane == 0) {
				const unsigned long long k0 = e[0];
				BucketHdr h;
				h.seg = (int32_t)((k0 >> 27) & 0x1FFFFFu);
				h.score = (int16_t)score;
				h.iscore = 0;
				h.first_seed = (uint16_t)(((k0 >> 11) & 0xFFFFu) + 1);
				h.nst = (uint16_t)min(na, SLOTS);
				h.eoff = hdr[b].eoff;
				hdr[b] = h;
			}
		}
		__syncthreads();
		// ---- index_score snapshot: score(seg) + score(seg-1) at the time of seg's last accept
		int nord = 0;
		for (int base = 0; base < nb; base += blockDim.x) {
			const int b = base + tid;
			int pass = 0;
			unsigned long long okey = 0;
			if (b < nb) {
				const int s0 = (int)sc.bstart[b], s1 = (int)sc.bstart[b + 1];
				const int seg = hdr[b].seg;
				int sk = hdr[b].score;
				if (b > 0 && hdr[b - 1].seg == seg - 1) {
					const uint32_t last_km = (uint32_t)((sc.ent[s1 - 1] >> 11) & 0xFFFFu);
					// accepted hits of seg-1 with k-mer ordinal <= last_km (same ordinal: seg-1 comes first)
					int lo = (int)sc.bstart[b - 1], hi = s0;
					const int p0 = lo;
					while (lo < hi) {
						const int mid = (lo + hi) >> 1;
						if ((uint32_t)((sc.ent[mid] >> 11) & 0xFFFFu) <= last_km) lo = mid + 1; else hi = mid;
					}
					if (lo > p0) sk += (int)(sc.ent[lo - 1] >> 48);
				}
				hdr[b].iscore = (int16_t)sk;
				pass = (int16_t)sk >= P.gate;
				okey = ((unsigned long long)hdr[b].first_seed << 48) | ((unsigned long long)(uint32_t)seg << 24) | (unsigned long long)(b & 0xFFFFFF);
			}
			int tot;
			const int ex = block_excl_scan(pass, &tot, misc);
			if (pass) sc.okeys[nord + ex] = okey;
			nord += tot;
		}
		__syncthreads();
		if (nord > 0) {
			int o2 = 1;
			while (o2 < nord) o2 <<= 1;
			for (int i = nord + tid; i < o2; i += blockDim.x) sc.okeys[i] = ~0ull;
			block_sort(sc.okeys, o2);
			for (int i = tid; i < nord; i += blockDim.x) ord[i] = (uint32_t)(sc.okeys[i] & 0xFFFFFFu);
		}
		D.nord = nord;
		if (tid == 0) P.desc[item] = D;
	}
}

// ------------------------------------------------------------------------------------------
struct WalkParams
{
	const unsigned char* arena;
	const StrandDesc* desc;
	int read0, nreads;
	const int2* qoffsz;
	int q_start_id;
	const int2* roffsz;     // index volume reads
	int r_nreads, r_start_id;
	int gate;               // 2 * min_kmer_match
	int min_span;           // min_kmer_dist (1800 for pacbio)
	int maxc;
	RawCand* lists;         // (2 * nreads) x maxc
	int32_t* nlist;         // 2 * nreads
};

// get_read_id_from_offset_list, split_database.cpp:16-35
__device__ int read_of_offset(const int2* __restrict__ a, int n, int offset)
{
	int left = 0, right = n - 1, mid = (left + right) / 2;
	if (a[right].x < offset) return right;
	while (left <= right) {
		const int2 o = a[mid];
		if (o.x <= offset && o.x + o.y > offset) return mid;
		if (o.x + o.y <= offset) left = mid + 1;
		else right = mid - 1;
		mid = (left + right) / 2;
	}
	return mid;
}

// index of the bucket holding `seg`, or -1
__device__ int find_bucket(const BucketHdr* hdr, int nb, int seg)
{
	int lo = 0, hi = nb - 1;
	while (lo <= hi) {
		const int mid = (lo + hi) >> 1;
		const int s = hdr[mid].seg;
		if (s == seg) return mid;
		if (s < seg) lo = mid + 1; else hi = mid - 1;
	}
	return -1;
}

constexpr int WALK_WARPS = 4;

__global__ void __launch_bounds__(WALK_WARPS * 32) k_walk(WalkParams P)
{
	__shared__ int s_loc[WALK_WARPS][2 * SLOTS + 16];
	__shared__ int s_seed[WALK_WARPS][2 * SLOTS + 16];
	__shared__ int s_score[WALK_WARPS][2 * SLOTS + 16];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int item = blockIdx.x * WALK_WARPS + warp;
	if (item >= 2 * P.nreads) return;
	int* w_loc = s_loc[warp];
	int* w_seed = s_seed[warp];
	int* w_score = s_score[warp];
	const StrandDesc D = P.desc[item];
	const int r = P.read0 + (item >> 1), chain = item & 1;
	const int qlen = P.qoffsz[r].y, qid = r + P.q_start_id;
	RawCand* list = P.lists + (size_t)item * P.maxc;
	int have = 0;
	// header scores / entries are private to this warp and mutated during the walk
	BucketHdr* hdr = (BucketHdr*)(P.arena + D.hdr_off);
	ushort2* ents = (ushort2*)(P.arena + D.ent_off);
	const uint32_t* ord = (const uint32_t*)(P.arena + D.ord_off);
	const int nb = D.nb;

	for (int oi = 0; oi < D.nord; ++oi) {
		const int b = (int)ord[oi];
		__syncwarp();
		const int seg = hdr[b].seg;
		const int cur = hdr[b].score;
		if (cur == 0) continue;
		int prev = 0;
		const bool has_prev = b > 0 && hdr[b - 1].seg == seg - 1;
		if (has_prev) prev = hdr[b - 1].score;
		const int origin = (prev > 0 ? seg - 1 : seg) * SEGW;
		// ---- window = previous bucket's seeds, then this bucket's (+2000)
		int n = 0;
		if (prev > 0) {
			const int pn = min(prev, SLOTS);
			const ushort2* pe = ents + hdr[b - 1].eoff;
			for (int j = lane; j < pn; j += 32) { const ushort2 e = pe[j]; w_loc[j] = e.x; w_seed[j] = (short)e.y; }
			n = pn;
		}
		{
			const int cn = min(cur, SLOTS);
			const ushort2* ce = ents + hdr[b].eoff;
			const int add = prev > 0 ? SEGW : 0;
			for (int j = lane; j < cn; j += 32) { const ushort2 e = ce[j]; w_loc[n + j] = e.x + add; w_seed[n + j] = (short)e.y; }
			n += cn;
		}
		for (int j = lane; j < n; j += 32) w_score[j] = 0;
		__syncwarp();
		// ---- find_location (pw_impl.cpp:161-239): mutual DDF scoring
		for (int i = lane; i < n - 1; i += 32) {
			const int li = w_loc[i], si = w_seed[i];
			int last_seed = si, mine = 0;
			for (int j = i + 1; j < n; ++j) {
				const int sj = w_seed[j], dl = w_loc[j] - li, ds = sj - si;
				if (last_seed != sj && ds > 0 && dl > 0 && dl < qlen && ddf_close(dl, ds)) {
					++mine;
					atomicAdd(&w_score[j], 1);
					last_seed = sj;
				}
			}
			if (mine) atomicAdd(&w_score[i], mine);
		}
		__syncwarp();
		int best = -1;
		for (int i = lane; i < n; i += 32) best = max(best, (w_score[i] << 7) | (127 - i));
		best = __reduce_max_sync(FULL, best);
		const int maxval = best >> 7, maxi = 127 - (best & 127);
		if (maxval < 5) continue;
		int rep = 0, lasti = 0;
		for (int i0 = 0; i0 < n; i0 += 32) {
			const int i = i0 + lane;
			const bool tie = i < n && i > maxi && w_score[i] == maxval;
			const unsigned m = __ballot_sync(FULL, tie);
			rep += __popc(m);
			if (m) lasti = i0 + 31 - __clz(m);
		}
		int anchor;
		if (rep == maxval) anchor = maxi;
		else {
			// first entry consistent with maxi (in index order, maxi included) whose position is non-zero,
			// else the last consistent one (the `loc[0]==0` test of :198,211,224)
			const int lm = w_loc[maxi], sm = w_seed[maxi];
			int first_nz = 0x7fffffff, last_any = -1;
			for (int i0 = 0; i0 < n; i0 += 32) {
				const int i = i0 + lane;
				bool f = false;
				if (i < n) {
					if (i < maxi) { const int dl = lm - w_loc[i], ds = sm - w_seed[i]; f = ds > 0 && dl > 0 && dl < qlen && ddf_close(dl, ds); }
					else if (i == maxi) f = true;
					else { const int dl = w_loc[i] - lm, ds = w_seed[i] - sm; f = ds > 0 && dl > 0 && dl <= qlen && ddf_close(dl, ds); }
				}
				const unsigned mf = __ballot_sync(FULL, f);
				const unsigned mz = __ballot_sync(FULL, f && w_loc[i < n ? i : 0] != 0);
				if (mf) last_any = i0 + 31 - __clz(mf);
				if (mz && first_nz == 0x7fffffff) first_nz = i0 + __ffs(mz) - 1;
			}
			anchor = first_nz != 0x7fffffff ? first_nz : last_any;
		}
		(void)lasti;
		const int ascore = w_score[anchor];
		if (ascore < P.gate + 2) continue;
		const int anchor_seed = w_seed[anchor];
		const int gpos = origin + w_loc[anchor];
		const int sidx = read_of_offset(P.roffsz, P.r_nreads, gpos);
		const int2 so = P.roffsz[sidx];
		const int sstart = so.x, ssize = so.y, send = sstart + ssize + 1;
		const int sid = sidx + P.r_start_id;
		if (sid > qid) continue;
		if (sid == qid) {
			// ---- purge the read's own span from the buckets (:371-383)
			int u = sstart / SEGW;
			{
				const int bi = find_bucket(hdr, nb, u);
				if (bi >= 0) {
					const int sc0 = hdr[bi].score, cut = sstart % SEGW;
					ushort2* e = ents + hdr[bi].eoff;
					const int m = min(sc0, SLOTS);
					int kept = 0;
					for (int j0 = 0; j0 < m; j0 += 32) {
						const int j = j0 + lane;
						unsigned short v = 0;
						bool keep = false;
						if (j < m) { v = e[j].x; keep = (int)v < cut; }
						const unsigned km = __ballot_sync(FULL, keep);
						__syncwarp();
						if (keep) e[kept + __popc(km & ((1u << lane) - 1u))].x = v;
						kept += __popc(km);
						__syncwarp();
					}
					if (lane == 0) hdr[bi].score = (int16_t)kept;
				}
			}
			++u;
			const int lastu = send / SEGW;
			{
				// buckets strictly inside the span lose everything
				int bi = -1;
				{
					int lo = 0, hi = nb;       // first bucket with seg >= u
					while (lo < hi) { const int mid = (lo + hi) >> 1; if (hdr[mid].seg < u) lo = mid + 1; else hi = mid; }
					bi = lo;
				}
				for (int x = bi + lane; x < nb && hdr[x].seg < lastu; x += 32) hdr[x].score = 0;
			}
			const int tail = max(u, lastu);
			{
				const int bi = find_bucket(hdr, nb, tail);
				if (bi >= 0) {
					__syncwarp();
					const int sc0 = hdr[bi].score, cut = send % SEGW;
					ushort2* e = ents + hdr[bi].eoff;
					const int m = min(sc0, SLOTS);
					int kept = 0;
					for (int j0 = 0; j0 < m; j0 += 32) {
						const int j = j0 + lane;
						unsigned short v = 0;
						bool keep = false;
						if (j < m) { v = e[j].x; keep = (int)v > cut; }
						const unsigned km = __ballot_sync(FULL, keep);
						__syncwarp();
						if (keep) e[kept + __popc(km & ((1u << lane) - 1u))].x = v;
						kept += __popc(km);
						__syncwarp();
					}
					if (lane == 0) hdr[bi].score = (int16_t)kept;
				}
			}
			__syncwarp();
			continue;
		}
		RawCand c;
		c.chain = chain;
		c.readno = sid; c.readstart = sstart;
		const int qoff = (anchor_seed - 1) * STRIDE;
		c.left1 = gpos - sstart + KMER - 1; c.right1 = send - gpos;
		c.left2 = qoff + KMER - 1; c.right2 = qlen - qoff;
		c.num1 = min(c.left1, c.left2); c.num2 = min(c.right1, c.right2);
		if (c.num1 + c.num2 < P.min_span) continue;
		c.loc1 = gpos - sstart; c.loc2 = qoff;
		// ---- neighbour votes (:405-438): every bucket of the overlap span, left then right
		int extra = 0;
		{
			const int nl = (c.num1 + SEGW - 1) / SEGW;
			const int lo_seg = max(0, seg - nl);
			for (int x = b - 1; x >= 0; --x) {
				const int u = hdr[x].seg;
				if (u < lo_seg) break;
				const int sc0 = hdr[x].score;
				if (sc0 <= 0) continue;
				const int cnt = min(sc0, SLOTS);
				const ushort2* e = ents + hdr[x].eoff;
				int ok = 0;
				for (int j0 = 0; j0 < cnt; j0 += 32) {
					const int j = j0 + lane;
					bool g = false;
					if (j < cnt) { const ushort2 v = e[j]; g = ddf_close(gpos - u * SEGW - (int)(short)v.x, anchor_seed - (int)(short)v.y); }
					ok += __popc(__ballot_sync(FULL, g));
				}
				extra += ok;
				if (ok * 1.0 / cnt > 0.4) { __syncwarp(); if (lane == 0) hdr[x].score = 0; }
			}
			const int nr = (c.num2 + SEGW - 1) / SEGW;
			const int hi_seg = seg + nr;
			for (int x = b + 1; x < nb; ++x) {
				const int u = hdr[x].seg;
				if (u > hi_seg) break;
				const int sc0 = hdr[x].score;
				if (sc0 <= 0) continue;
				const int cnt = min(sc0, SLOTS);
				const ushort2* e = ents + hdr[x].eoff;
				int ok = 0;
				for (int j0 = 0; j0 < cnt; j0 += 32) {
					const int j = j0 + lane;
					bool g = false;
					if (j < cnt) { const ushort2 v = e[j]; g = ddf_close(u * SEGW + (int)(short)v.x - gpos, (int)(short)v.y - anchor_seed); }
					ok += __popc(__ballot_sync(FULL, g));
				}
				extra += ok;
				if (ok * 1.0 / cnt > 0.4) { __syncwarp(); if (lane == 0) hdr[x].score = 0; }
			}
		}
		c.score = ascore + extra;
		// ---- stable insertion into the score-descending list capped at maxc (:442-455)
		__syncwarp();
		int at = 0;
		{
			int cntge = 0;
			for (int i0 = 0; i0 < have; i0 += 32) {
				const int i = i0 + lane;
				cntge += __popc(__ballot_sync(FULL, i < have && list[i].score >= c.score));
			}
			at = cntge;
		}
		if (at < P.maxc) {
			const int last = min(have, P.maxc - 1);       // new index of the last surviving element
			// shift [at, last) up by one, highest first, 32 at a time
			for (int hi = last; hi > at; hi -= 32) {
				const int i = hi - lane;                  // destination index
				RawCand tmp;
				const bool mv = i > at;
				if (mv) tmp = list[i - 1];
				__syncwarp();
				if (mv) list[i] = tmp;
				__syncwarp();
			}
			if (lane == 0) list[at] = c;
		}
		if (have < P.maxc) ++have;
		__syncwarp();
	}
	if (lane == 0) P.nlist[item] = have;
}

// stable merge of the F list and the R list of a read (pw_impl.cpp:751-765 share one list)
__global__ void k_merge(const RawCand* __restrict__ lists, const int32_t* __restrict__ nlist, int read0, int nreads,
                        int maxc, RawCand* __restrict__ out, int32_t* __restrict__ counts)
{
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (warp >= nreads) return;
	const RawCand* F = lists + (size_t)(2 * warp) * maxc;
	const RawCand* R = F + maxc;
	const int nf = nlist[2 * warp], nr = nlist[2 * warp + 1];
	RawCand* o = out + (size_t)(read0 + warp) * maxc;
	for (int i = lane; i < nf; i += 32) {
		const int s = F[i].score;
		int lo = 0, hi = nr;                 // # R with score > s
		while (lo < hi) { const int mid = (lo + hi) >> 1; if (R[mid].score > s) lo = mid + 1; else hi = mid; }
		const int rank = i + lo;
		if (rank < maxc) o[rank] = F[i];
	}
	for (int j = lane; j < nr; j += 32) {
		const int s = R[j].score;
		int lo = 0, hi = nf;                 // # F with score >= s
		while (lo < hi) { const int mid = (lo + hi) >> 1; if (F[mid].score >= s) lo = mid + 1; else hi = mid; }
		const int rank = j + lo;
		if (rank < maxc) o[rank] = R[j];
	}
	if (lane == 0) counts[read0 + warp] = min(nf + nr, maxc);
}

}  // namespace

int seed_candidates(Ctx* c, const DIndex* idx, const DVolume* ref, const DVolume* reads, const mecat_pw_params* p,
                    int read_begin, int read_end, RawCand* d_cands, int32_t* d_counts)
{
	// reads outside [read_begin, read_end) get no candidates (d_counts is cleared by the caller)
	const int N = read_end;
	if (read_end <= read_begin) return 0;
	const int maxc = p->num_candidates;
	const int max_nk = reads->max_read >= KMER ? (reads->max_read - KMER) / STRIDE + 1 : 1;
	if (max_nk > MAX_KM) MB_FAIL(c, "seeding: reads longer than %d bp are outside this path (seed ordinals are 16-bit in the reference)", MAX_KM * STRIDE);
	const int kcap = max_nk + 32;
	// per-strand capacity for the collected hits of wanted buckets: the true upper bound, every sampled k-mer with a full
	// list (MAX_OCC positions) -- a strand of a repeat-rich volume gets close to it (15 kb read: 1 500 x 128 = 192 000
	// hits, 7 MB of scratch per CTA), and the reference maps such reads like any other
	int hcap = 1 << 16;
	while ((long long)hcap < (long long)MAX_OCC * max_nk && hcap < (1 << 23)) hcap <<= 1;
	// MECAT_B200_SEED_CTAS = 1 | 2: shape of k_seed (see SeedShape)
	int seed_ctas = 2;
	if (const char* e = getenv("MECAT_B200_SEED_CTAS")) seed_ctas = atoi(e) == 1 ? 1 : 2;
	const int nctas = c->sm_count * seed_ctas;
	const int seed_threads = 1024 / seed_ctas;
	const int cnt_words = seed_ctas == 1 ? CNT_SLOTS / 2 : CNT_SLOTS / 4;
	int batch = 8192;
	if (batch > read_end - read_begin) batch = read_end - read_begin;
	unsigned long long arena_bytes = 1ull << 30;

	const size_t smem = (size_t)(cnt_words + BIT_WORDS + KCACHE) * 4 + KCACHE + 64 * 4 + (size_t)(seed_threads / 32) * 123 * 4;
	MB_CUDA(c, cudaFuncSetAttribute(k_seed<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	MB_CUDA(c, cudaFuncSetAttribute(k_seed<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

	unsigned char* d_arena = nullptr;
	StrandDesc* d_desc = nullptr;
	SeedScratch* d_scratch = nullptr;
	unsigned char* d_pool = nullptr;
	RawCand* d_lists = nullptr;
	int32_t* d_nlist = nullptr;
	std::vector<StrandDesc> h_desc;
	auto body = [&]() -> int {
		// per-CTA scratch
		const size_t per = (((size_t)kcap * 4 + 255) & ~255ull) + (((size_t)kcap + 255) & ~255ull) + (size_t)hcap * 8 * 3 + (size_t)(hcap + 64) * 4;
		MB_CUDA(c, c->dmalloc((void**)&d_pool, per * nctas));
		std::vector<SeedScratch> hs(nctas);
		for (int i = 0; i < nctas; ++i) {
			unsigned char* b = d_pool + per * i;
			hs[i].kb = (uint32_t*)b; b += ((size_t)kcap * 4 + 255) & ~255ull;
			hs[i].kc = (uint8_t*)b; b += ((size_t)kcap + 255) & ~255ull;
			hs[i].keys = (unsigned long long*)b; b += (size_t)hcap * 8;
			hs[i].ent = (unsigned long long*)b; b += (size_t)hcap * 8;
			hs[i].okeys = (unsigned long long*)b; b += (size_t)hcap * 8;
			hs[i].bstart = (uint32_t*)b;
		}
		MB_CUDA(c, c->alloc(&d_scratch, (size_t)nctas));
		MB_CUDA(c, cudaMemcpyAsync(d_scratch, hs.data(), sizeof(SeedScratch) * nctas, cudaMemcpyHostToDevice, c->stream));
		MB_CUDA(c, c->dmalloc((void**)&d_arena, arena_bytes));
		MB_CUDA(c, c->alloc(&d_desc, 2 * (size_t)batch));
		MB_CUDA(c, c->alloc(&d_lists, 2 * (size_t)batch * maxc));
		MB_CUDA(c, c->alloc(&d_nlist, 2 * (size_t)batch));
		unsigned long long* d_cursor = c->d_counters + 1;
		unsigned int* d_work = (unsigned int*)(c->d_counters + 2);
		unsigned long long* d_hits = c->d_counters + 3;
		MB_CUDA(c, cudaMemsetAsync(d_hits, 0, 8, c->stream));
		h_desc.resize(2 * (size_t)batch);
		for (int r0 = read_begin; r0 < N;) {
			const int nb = std::min(batch, N - r0);
			MB_CUDA(c, cudaMemsetAsync(c->d_counters + 1, 0, 16, c->stream));   // cursor + work counter
			SeedParams S;
			S.begin = idx->begin; S.pos = idx->pos;
			S.qfwd = reads->fwd; S.qrev = reads->rev; S.qoffsz = reads->offsz; S.qN = reads->num_bases;
			S.read0 = r0; S.nreads = nb; S.gate = 2 * p->min_kmer_match;
			S.arena = d_arena; S.arena_bytes = arena_bytes; S.arena_cursor = d_cursor; S.work_counter = d_work;
			S.desc = d_desc; S.scratch = d_scratch; S.kcap = kcap; S.hcap = hcap; S.hit_counter = d_hits;
			{
				KScope ks(c, MECAT_K_SEED);
				if (seed_ctas == 1) k_seed<1><<<nctas, seed_threads, smem, c->stream>>>(S);
				else k_seed<2><<<nctas, seed_threads, smem, c->stream>>>(S);
			}
			MB_CUDA(c, cudaGetLastError());
			// status check (arena overflow -> retry the batch with fewer reads)
			MB_CUDA(c, cudaMemcpyAsync(h_desc.data(), d_desc, sizeof(StrandDesc) * 2 * (size_t)nb, cudaMemcpyDeviceToHost, c->stream));
			MB_CUDA(c, cudaStreamSynchronize(c->stream));
			bool overflow = false;
			for (int i = 0; i < 2 * nb; ++i) {
				if (h_desc[i].status == 1) overflow = true;
				else if (h_desc[i].status == 2) MB_FAIL(c, "seeding: read %d is too long for this path", r0 + i / 2);
				else if (h_desc[i].status == 3) MB_FAIL(c, "seeding: read %d has more candidate seeds than the per-strand capacity (%d)", r0 + i / 2, hcap);
			}
			if (overflow) {
				if (nb == 1) MB_FAIL(c, "seeding: one read overflows the bucket arena");
				batch = std::max(1, nb / 2);
				continue;
			}
			WalkParams Wp;
			Wp.arena = d_arena; Wp.desc = d_desc; Wp.read0 = r0; Wp.nreads = nb;
			Wp.qoffsz = reads->offsz; Wp.q_start_id = reads->start_read_id;
			Wp.roffsz = ref->offsz; Wp.r_nreads = ref->num_reads; Wp.r_start_id = ref->start_read_id;
			Wp.gate = 2 * p->min_kmer_match; Wp.min_span = p->tech == 1 ? 400 : 1800 /* min_kmer_dist, pw_impl.cpp:843-849 */; Wp.maxc = maxc;
			Wp.lists = d_lists; Wp.nlist = d_nlist;
			{
				KScope ks(c, MECAT_K_WALK);
				k_walk<<<(2 * nb + WALK_WARPS - 1) / WALK_WARPS, WALK_WARPS * 32, 0, c->stream>>>(Wp);
			}
			{
				KScope ks(c, MECAT_K_MERGE);
				k_merge<<<(nb * 32 + 127) / 128, 128, 0, c->stream>>>(d_lists, d_nlist, r0, nb, maxc, d_cands, d_counts);
			}
			MB_CUDA(c, cudaGetLastError());
			r0 += nb;
		}
		unsigned long long hits = 0;
		MB_CUDA(c, cudaMemcpyAsync(&hits, d_hits, 8, cudaMemcpyDeviceToHost, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		c->resolve_timers();
		c->stats.num_hits += (int64_t)hits;
		return 0;
	};
	int rc = body();
	c->dfree(d_arena); c->dfree(d_desc); c->dfree(d_scratch); c->dfree(d_pool); c->dfree(d_lists); c->dfree(d_nlist);
	return rc;
}

}  // namespace mb
