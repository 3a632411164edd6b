// mecat_b200/csrc/volume.cu -- packed read volumes resident in HBM.
//
// Replaces load_volume / extract_one_seq / reverse_complement of the reference
// (src/common/split_database.cpp:122-133,156-181; src/mecat2pw/pw_impl.cpp:69-81): instead
// of unpacking every read to one byte per base and building its reverse complement on the
// CPU for every query, the 2-bit volume is copied to the device once and re-laid out in two
// orientations (forward, reversed) from which every strand / walking direction of the
// path is a plain forward walk (complement = bitwise NOT).
#include "common.cuh"

namespace mb {

// pac byte j holds bases 4j..4j+3 with base 4j in bits 7..6 (packed_db.h:98-107).
// A little-endian 32-bit load of 4 such bytes only needs the four 2-bit groups of every
// byte swapped end for end to become "base i at bits 2*(i%16)".
__global__ void k_orient_fwd(const uint32_t* __restrict__ pac32, uint32_t* __restrict__ fwd, size_t src_words,
                             size_t words)
{
	size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if (w >= words) return;
	uint32_t x = w < src_words ? pac32[w] : 0u;
	x = ((x & 0x03030303u) << 6) | ((x & 0x0C0C0C0Cu) << 2) | ((x & 0x30303030u) >> 2) | ((x & 0xC0C0C0C0u) >> 6);
	fwd[w] = x;
}

// rev base i = base N-1-i
__global__ void k_orient_rev(const uint32_t* __restrict__ fwd, uint32_t* __restrict__ rev, int64_t nbases,
                             size_t words)
{
	size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if (w >= words) return;
	int64_t s = nbases - 16 - 16 * (int64_t)w;   // first forward base covered by this word
	uint32_t x;
	if (s >= 0) x = ld_bases32(fwd, (uint32_t)s);
	else if (s > -16) x = __ldg(fwd) << (2 * (int)(-s));
	else x = 0u;
	rev[w] = rev_groups2(x);
}

// Shared tail of both constructors: offsets to the device, packed bytes (already on the device at
// d_pac, readable as 32-bit words, zero padded to src_words) re-laid out in both orientations.
static int volume_build(Ctx* c, int num_reads, int num_bases, int start_read_id, const int32_t* h_offsz,
                        const uint32_t* d_pac, size_t src_words, DVolume** out)
{
	DVolume* d = new DVolume;
	d->num_reads = num_reads;
	d->num_bases = num_bases;
	d->start_read_id = start_read_id;
	d->h_offsz.assign(h_offsz, h_offsz + 2 * (size_t)num_reads);
	for (int i = 0; i < num_reads; ++i) {
		int off = h_offsz[2 * i], sz = h_offsz[2 * i + 1];
		if (off < 0 || sz < 0 || (int64_t)off + sz > num_bases) { delete d; MB_FAIL(c, "volume: read %d out of range", i); }
		if (sz > d->max_read) d->max_read = sz;
	}
	d->words = ((size_t)num_bases + 15) / 16 + 8;
	auto fail = [&](cudaError_t e, const char* what) {
		char b[256];
		snprintf(b, sizeof b, "volume: %s: %s", what, cudaGetErrorString(e));
		c->err = b;
		volume_release(c, d);
		return 1;
	};
	cudaError_t e;
	if ((e = c->dmalloc((void**)&d->fwd, d->words * 4)) != cudaSuccess) return fail(e, "cudaMalloc fwd");
	if ((e = c->dmalloc((void**)&d->rev, d->words * 4)) != cudaSuccess) return fail(e, "cudaMalloc rev");
	if ((e = c->dmalloc((void**)&d->offsz, sizeof(int2) * (size_t)(num_reads ? num_reads : 1))) != cudaSuccess) return fail(e, "cudaMalloc offsets");
	if (num_reads && (e = cudaMemcpyAsync(d->offsz, h_offsz, sizeof(int2) * (size_t)num_reads, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "H2D offsets");
	const int T = 256;
	const unsigned G = (unsigned)((d->words + T - 1) / T);
	{
		KScope ks(c, MECAT_K_ORIENT, 2);
		k_orient_fwd<<<G, T, 0, c->stream>>>(d_pac, d->fwd, src_words, d->words);
		k_orient_rev<<<G, T, 0, c->stream>>>(d->fwd, d->rev, num_bases, d->words);
	}
	if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "orient kernels");
	c->resolve_timers();
	c->stats.h2d_bytes += (int64_t)sizeof(int2) * num_reads;
	*out = d;
	return 0;
}

int volume_upload(Ctx* c, const mecat_volume* v, DVolume** out)
{
	if (!v || v->num_reads < 0 || v->num_bases < 0) MB_FAIL(c, "volume_upload: bad volume");
	const size_t pac_bytes = ((size_t)v->num_bases + 3) / 4;
	const size_t src_words = (pac_bytes + 3) / 4;
	uint32_t* d_pac = nullptr;
	MB_CUDA(c, c->dmalloc((void**)&d_pac, src_words * 4 + 8));      // an empty volume still gets its 8 zero bytes
	cudaError_t e = cudaMemsetAsync(d_pac + (src_words ? src_words - 1 : 0), 0, 8, c->stream);
	if (e == cudaSuccess && pac_bytes) e = cudaMemcpyAsync(d_pac, v->pac, pac_bytes, cudaMemcpyHostToDevice, c->stream);
	if (e != cudaSuccess) { c->dfree(d_pac); MB_FAIL(c, "volume_upload: H2D: %s", cudaGetErrorString(e)); }
	int rc = volume_build(c, v->num_reads, v->num_bases, v->start_read_id, v->offset_size, d_pac, src_words, out);
	c->dfree(d_pac);
	if (!rc) c->stats.h2d_bytes += (int64_t)pac_bytes;
	return rc;
}

// The packed bytes are already in device memory (e.g. received from a peer GPU over NCCL).
// d_pac must be 4-byte aligned and readable up to the next multiple of 4 bytes.
int volume_from_device(Ctx* c, int num_reads, int num_bases, int start_read_id, const int32_t* h_offsz,
                       const uint8_t* d_pac, DVolume** out)
{
	if (num_reads < 0 || num_bases < 0 || ((uintptr_t)d_pac & 3)) MB_FAIL(c, "volume_from_device: bad arguments");
	const size_t pac_bytes = ((size_t)num_bases + 3) / 4;
	const size_t src_words = pac_bytes / 4;      // whole words only; a ragged tail is copied aside below
	const size_t tail = pac_bytes - src_words * 4;
	if (tail == 0) return volume_build(c, num_reads, num_bases, start_read_id, h_offsz, (const uint32_t*)d_pac, src_words, out);
	uint32_t* tmp = nullptr;
	MB_CUDA(c, c->dmalloc((void**)&tmp, (src_words + 2) * 4));
	cudaError_t e = cudaMemsetAsync(tmp + src_words, 0, 8, c->stream);
	if (e == cudaSuccess) e = cudaMemcpyAsync(tmp, d_pac, pac_bytes, cudaMemcpyDeviceToDevice, c->stream);
	if (e != cudaSuccess) { c->dfree(tmp); MB_FAIL(c, "volume_from_device: copy: %s", cudaGetErrorString(e)); }
	int rc = volume_build(c, num_reads, num_bases, start_read_id, h_offsz, tmp, src_words + 1, out);
	c->dfree(tmp);
	return rc;
}

// ---- 2-bit packing on the device: add_one_seq / PackedDB::set_char for every read of a volume
// (src/common/split_database.cpp:103-119, src/common/packed_db.h:98-101).  The host has found the records (where each
// read's letters start in the text, and the volume offsets the reference's splitting rule gives them); one warp packs one
// read, a lane 16 letters at a time into one 32-bit word of `pac`: letter -> code through get_dna_encode_table
// (src/common/defs.cpp:3-36), code << shift OR-ed into its byte and cut to the byte, so that a code above 3 (N = 14, any
// other character 16) spills into the bases packed before it in the same byte exactly like set_char does.  The first and
// last word of a read are shared with its neighbours and are OR-ed in atomically.
struct PackRead { int64_t src; int32_t dst, len; };

__global__ void __launch_bounds__(256) k_pack_text(const unsigned char* __restrict__ text, const PackRead* __restrict__ reads, int n,
                                                   uint32_t* __restrict__ pac32)
{
	__shared__ uint8_t enc[256];
	{
		// get_dna_encode_table: IUPAC letters in either case, '-' = 15, everything else 16
		const int t = threadIdx.x;
		uint8_t v = 16;
		const int u = t >= 'a' && t <= 'z' ? t - 32 : t;
		switch (u) {
		case 'A': v = 0; break; case 'C': v = 1; break; case 'G': v = 2; break; case 'T': v = 3; break;
		case 'R': v = 4; break; case 'Y': v = 5; break; case 'M': v = 6; break; case 'K': v = 7; break;
		case 'W': v = 8; break; case 'S': v = 9; break; case 'B': v = 10; break; case 'D': v = 11; break;
		case 'H': v = 12; break; case 'V': v = 13; break; case 'N': v = 14; break; case '-': v = 15; break;
		}
		enc[t] = v;
	}
	__syncthreads();
	const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (warp >= n) return;
	const PackRead r = reads[warp];
	if (r.len <= 0) return;
	const unsigned char* src = text + r.src;
	const uint32_t w0 = (uint32_t)r.dst >> 4, w1 = ((uint32_t)r.dst + (uint32_t)r.len - 1u) >> 4;
	for (uint32_t w = w0 + (uint32_t)lane; w <= w1; w += 32u) {
		const int64_t first = (int64_t)w * 16 - (int64_t)r.dst;            // read coordinate of the word's first base
		uint32_t x = 0;
		#pragma unroll
		for (int k = 0; k < 16; ++k) {
			const int64_t i = first + k;
			if (i < 0 || i >= r.len) continue;
			const uint32_t code = enc[src[i]];
			// base 16 w + k lives in byte k / 4 of the word at shift ((~k) & 3) * 2; the OR is cut to that byte
			x |= ((code << (((~k) & 3) << 1)) & 0xFFu) << ((k >> 2) << 3);
		}
		if (w == w0 || w == w1) atomicOr(pac32 + w, x);
		else pac32[w] = x;
	}
}

int volume_from_text(Ctx* c, const char* text, size_t text_bytes, const int64_t* src_off, const int32_t* h_offsz, int num_reads,
                     int num_bases, int start_read_id, uint8_t* pac_out, DVolume** out)
{
	if (num_reads < 0 || num_bases < 0 || (num_reads && (!text || !src_off || !h_offsz))) MB_FAIL(c, "volume_from_text: bad arguments");
	std::vector<PackRead> pr((size_t)num_reads);
	for (int i = 0; i < num_reads; ++i) {
		const int32_t off = h_offsz[2 * i], len = h_offsz[2 * i + 1];
		if (off < 0 || len < 0 || (int64_t)off + len > num_bases || src_off[i] < 0 || (uint64_t)src_off[i] + (uint64_t)len > text_bytes)
			MB_FAIL(c, "volume_from_text: read %d out of range", i);
		pr[(size_t)i].src = src_off[i]; pr[(size_t)i].dst = off; pr[(size_t)i].len = len;
	}
	const size_t pac_bytes = ((size_t)num_bases + 3) / 4, words = (pac_bytes + 3) / 4 + 2;
	unsigned char* d_text = nullptr;
	PackRead* d_pr = nullptr;
	uint32_t* d_pac = nullptr;
	auto body = [&]() -> int {
		MB_CUDA(c, c->dmalloc((void**)&d_text, text_bytes + 16));
		MB_CUDA(c, c->alloc(&d_pr, (size_t)(num_reads ? num_reads : 1)));
		MB_CUDA(c, c->alloc(&d_pac, words));
		MB_CUDA(c, cudaMemsetAsync(d_pac, 0, words * 4, c->stream));
		if (text_bytes) MB_CUDA(c, cudaMemcpyAsync(d_text, text, text_bytes, cudaMemcpyHostToDevice, c->stream));
		if (num_reads) MB_CUDA(c, cudaMemcpyAsync(d_pr, pr.data(), sizeof(PackRead) * (size_t)num_reads, cudaMemcpyHostToDevice, c->stream));
		if (num_reads) {
			KScope ks(c, MECAT_K_ORIENT);
			k_pack_text<<<(unsigned)(((size_t)num_reads * 32 + 255) / 256), 256, 0, c->stream>>>(d_text, d_pr, num_reads, d_pac);
		}
		MB_CUDA(c, cudaGetLastError());
		if (pac_out && pac_bytes) MB_CUDA(c, cudaMemcpyAsync(pac_out, d_pac, pac_bytes, cudaMemcpyDeviceToHost, c->stream));
		c->stats.h2d_bytes += (int64_t)(text_bytes + sizeof(PackRead) * (size_t)num_reads);
		if (pac_out) c->stats.d2h_bytes += (int64_t)pac_bytes;
		return volume_build(c, num_reads, num_bases, start_read_id, h_offsz, d_pac, words - 1, out);     // synchronises the stream
	};
	const int rc = body();
	c->dfree(d_text); c->dfree(d_pr); c->dfree(d_pac);
	return rc;
}

// ---- a working volume gathered from reads of several resident volumes (mecat2cns on read sets larger than one volume)
struct GatherRead
{
	const uint32_t* src;      // forward words of the source volume
	uint32_t src_off;         // first base of the read there
	uint32_t dst_off;         // first base of the read in the working volume
	int32_t len;
};

// One warp per read: destination word j of the read's span takes 16 bases from the source at the matching offset.  The
// first and the last word of a span are shared with the neighbouring reads (and the pad base between them, which stays
// 0 like in PackedDB), so they are OR-ed into a zeroed array; the words in between are plain stores.
__global__ void __launch_bounds__(256) k_gather_reads(const GatherRead* __restrict__ reads, int n, uint32_t* __restrict__ fwd)
{
	const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (warp >= n) return;
	const GatherRead r = reads[warp];
	if (r.len <= 0) return;
	const uint32_t w0 = r.dst_off >> 4, w1 = (r.dst_off + (uint32_t)r.len - 1u) >> 4;
	for (uint32_t w = w0 + (uint32_t)lane; w <= w1; w += 32u) {
		// bases [16 w, 16 w + 16) of the working volume, clipped to the read
		const int64_t first = (int64_t)w * 16 - (int64_t)r.dst_off;          // read coordinate of the word's first base (may be < 0)
		const int lo = first < 0 ? (int)-first : 0;                          // bases of the word before the read
		const int64_t left = (int64_t)r.len - first;
		const int hi = left < 16 ? (int)left : 16;                          // bases of the word inside the read end at `hi`
		uint32_t x = ld_bases32(r.src, (uint32_t)((int64_t)r.src_off + first + lo)) << (2 * lo);
		if (hi < 16) x &= (1u << (2 * hi)) - 1u;
		if (lo > 0) x &= ~((1u << (2 * lo)) - 1u);
		if (w == w0 || w == w1) atomicOr(fwd + w, x);
		else fwd[w] = x;
	}
}

int volume_gather(Ctx* c, const DVolume* const* src, const int32_t* h_src_vol, const int32_t* h_src_read, int n, DVolume** out)
{
	DVolume* d = new DVolume;
	std::vector<GatherRead> g((size_t)n);
	d->h_offsz.resize(2 * (size_t)n);
	int64_t curr = 0;
	for (int i = 0; i < n; ++i) {
		const DVolume* S = src[h_src_vol[i]];
		const int32_t off = S->h_offsz[2 * (size_t)h_src_read[i]], len = S->h_offsz[2 * (size_t)h_src_read[i] + 1];
		if (curr + len + 1 > 0x7fffffffLL) { delete d; MB_FAIL(c, "volume_gather: the working set of reads exceeds one volume"); }
		g[(size_t)i].src = S->fwd; g[(size_t)i].src_off = (uint32_t)off; g[(size_t)i].dst_off = (uint32_t)curr; g[(size_t)i].len = len;
		d->h_offsz[2 * (size_t)i] = (int32_t)curr; d->h_offsz[2 * (size_t)i + 1] = len;
		if (len > d->max_read) d->max_read = len;
		curr += len + 1;                   // pad base, split_database.cpp:251
	}
	d->num_reads = n; d->num_bases = (int32_t)curr; d->start_read_id = 0;
	d->words = ((size_t)curr + 15) / 16 + 8;
	GatherRead* d_g = nullptr;
	auto fail = [&](cudaError_t e, const char* what) {
		char b[256];
		snprintf(b, sizeof b, "volume_gather: %s: %s", what, cudaGetErrorString(e));
		c->err = b;
		c->dfree(d_g);
		volume_release(c, d);
		return 1;
	};
	cudaError_t e;
	if ((e = c->dmalloc((void**)&d->fwd, d->words * 4)) != cudaSuccess) return fail(e, "cudaMalloc fwd");
	if ((e = c->dmalloc((void**)&d->rev, d->words * 4)) != cudaSuccess) return fail(e, "cudaMalloc rev");
	if ((e = c->dmalloc((void**)&d->offsz, sizeof(int2) * (size_t)(n ? n : 1))) != cudaSuccess) return fail(e, "cudaMalloc offsets");
	if ((e = c->dmalloc((void**)&d_g, sizeof(GatherRead) * (size_t)(n ? n : 1))) != cudaSuccess) return fail(e, "cudaMalloc gather list");
	if ((e = cudaMemsetAsync(d->fwd, 0, d->words * 4, c->stream)) != cudaSuccess) return fail(e, "memset");
	if (n) {
		if ((e = cudaMemcpyAsync(d->offsz, d->h_offsz.data(), sizeof(int2) * (size_t)n, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "H2D offsets");
		if ((e = cudaMemcpyAsync(d_g, g.data(), sizeof(GatherRead) * (size_t)n, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "H2D gather list");
	}
	{
		KScope ks(c, MECAT_K_ORIENT, 2);
		if (n) k_gather_reads<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, c->stream>>>(d_g, n, d->fwd);
		k_orient_rev<<<(unsigned)((d->words + 255) / 256), 256, 0, c->stream>>>(d->fwd, d->rev, curr, d->words);
	}
	if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "gather kernels");
	c->resolve_timers();
	c->dfree(d_g);
	c->stats.h2d_bytes += (int64_t)(sizeof(int2) + sizeof(GatherRead)) * n;
	*out = d;
	return 0;
}

void volume_release(Ctx* c, DVolume* v)
{
	if (!v) return;
	c->dfree(v->offsz);
	c->dfree(v->fwd);
	c->dfree(v->rev);
	delete v;
}

}  // namespace mb
