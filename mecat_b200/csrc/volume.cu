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

int volume_upload(Ctx* c, const mecat_volume* v, DVolume** out)
{
	if (!v || v->num_reads < 0 || v->num_bases < 0) MB_FAIL(c, "volume_upload: bad volume");
	DVolume* d = new DVolume;
	d->num_reads = v->num_reads;
	d->num_bases = v->num_bases;
	d->start_read_id = v->start_read_id;
	d->h_offsz.assign(v->offset_size, v->offset_size + 2 * (size_t)v->num_reads);
	for (int i = 0; i < v->num_reads; ++i) {
		int off = v->offset_size[2 * i], sz = v->offset_size[2 * i + 1];
		if (off < 0 || sz < 0 || (int64_t)off + sz > v->num_bases) { delete d; MB_FAIL(c, "volume_upload: read %d out of range", i); }
		if (sz > d->max_read) d->max_read = sz;
	}
	const size_t pac_bytes = ((size_t)v->num_bases + 3) / 4;
	const size_t src_words = (pac_bytes + 3) / 4;
	d->words = ((size_t)v->num_bases + 15) / 16 + 8;
	uint32_t* d_pac = nullptr;
	auto fail = [&](cudaError_t e, const char* what) {
		char b[256];
		snprintf(b, sizeof b, "volume_upload: %s: %s", what, cudaGetErrorString(e));
		c->err = b;
		c->dfree(d_pac);
		volume_release(c, d);
		return 1;
	};
	cudaError_t e;
	if ((e = c->dmalloc((void**)&d_pac, src_words * 4 + 4)) != cudaSuccess) return fail(e, "cudaMalloc pac");
	if ((e = c->dmalloc((void**)&d->fwd, d->words * 4)) != cudaSuccess) return fail(e, "cudaMalloc fwd");
	if ((e = c->dmalloc((void**)&d->rev, d->words * 4)) != cudaSuccess) return fail(e, "cudaMalloc rev");
	if ((e = c->dmalloc((void**)&d->offsz, sizeof(int2) * (size_t)(v->num_reads ? v->num_reads : 1))) != cudaSuccess) return fail(e, "cudaMalloc offsets");
	if ((e = cudaMemsetAsync(d_pac, 0, src_words * 4 + 4, c->stream)) != cudaSuccess) return fail(e, "memset");
	if (pac_bytes && (e = cudaMemcpyAsync(d_pac, v->pac, pac_bytes, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "H2D pac");
	if (v->num_reads && (e = cudaMemcpyAsync(d->offsz, v->offset_size, sizeof(int2) * (size_t)v->num_reads, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return fail(e, "H2D offsets");
	const int T = 256;
	const unsigned G = (unsigned)((d->words + T - 1) / T);
	{
		KScope ks(c, MECAT_K_ORIENT, 2);
		k_orient_fwd<<<G, T, 0, c->stream>>>(d_pac, d->fwd, src_words, d->words);
		k_orient_rev<<<G, T, 0, c->stream>>>(d->fwd, d->rev, v->num_bases, d->words);
	}
	if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e, "orient kernels");
	c->resolve_timers();
	c->stats.h2d_bytes += (int64_t)pac_bytes + (int64_t)sizeof(int2) * v->num_reads;
	c->dfree(d_pac);
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
