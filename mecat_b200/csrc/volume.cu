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

void volume_release(Ctx* c, DVolume* v)
{
	if (!v) return;
	c->dfree(v->offsz);
	c->dfree(v->fwd);
	c->dfree(v->rev);
	delete v;
}

}  // namespace mb
