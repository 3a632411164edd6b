// tests/xdrop_host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the product's nanopore extension body (mecat_b200/csrc/xdrop_core.cuh: block DP, packed trace-back, chain of
// blocks, the assembly of XdropAligner::go) on the host, one "thread" at a time over the same packed-word walks the
// CUDA kernel of xdrop.cu hands it, so that the CPU test-suite can compare the statements the GPU executes with the
// oracle and with the unmodified XdropAligner.  Compiled by tests/util.py; never part of the product library.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../mecat_b200/csrc/xdrop_core.cuh"

namespace {

// the device layout of volume.cu: base p at bits 2 (p % 16) of word p / 16
std::vector<uint32_t> words_of(const char* codes, int n, bool reversed)
{
	std::vector<uint32_t> w((size_t)(n / 16 + 4), 0u);
	for (int p = 0; p < n; ++p) {
		const uint32_t b = (uint32_t)(reversed ? codes[n - 1 - p] : codes[p]) & 3u;
		w[(size_t)(p >> 4)] |= b << ((p & 15) << 1);
	}
	return w;
}

}  // namespace

extern "C" int xh_go(const char* q, int qstart, int qsize, const char* t, int tstart, int tsize, int min_aln, int32_t* out,
                     char* qstr, char* tstr, int cap, int want_cols)
{
	using namespace mbx;
	const std::vector<uint32_t> qf = words_of(q, qsize, false), qr = words_of(q, qsize, true);
	const std::vector<uint32_t> tf = words_of(t, tsize, false), tr = words_of(t, tsize, true);
	std::vector<unsigned char> mem(SCRATCH_BYTES + 64, 0xAB);      // like device memory: never zero by luck
	Scratch S;
	unsigned char* p = mem.data();
	S.sc = (Cell*)p; p += sizeof(Cell) * SC_CELLS;
	S.tb = (uint32_t*)p; p += 4 * (size_t)TB_WORDS;
	S.row_first = (int32_t*)p; p += 4 * (size_t)ROWS;
	S.row_word = (int32_t*)p;
	std::vector<RingCell> ring((size_t)RING * 3);          // stride 3: like the interleaved shared-memory layout of the kernel
	S.ring = (want_cols & 2) ? nullptr : ring.data(); S.ring_stride = 3;      // bit 1 of want_cols: the global row only
	want_cols &= 1;
	Half H[2];
	std::vector<char> cq[2], ct[2];
	for (int right = 0; right < 2; ++right) {
		Seq Q, T;
		if (right) {
			Q.arr = qf.data(); Q.g0 = (uint32_t)qstart; Q.comp = 0; Q.len = qsize - qstart;
			T.arr = tf.data(); T.g0 = (uint32_t)tstart; T.comp = 0; T.len = tsize - tstart;
		} else {
			Q.arr = qr.data(); Q.g0 = (uint32_t)(qsize - qstart); Q.comp = 0; Q.len = qstart;
			T.arr = tr.data(); T.g0 = (uint32_t)(tsize - tstart); T.comp = 0; T.len = tstart;
		}
		const int slot = Q.len + T.len + 8;
		cq[right].assign((size_t)slot, '?'); ct[right].assign((size_t)slot, '?');
		if (want_cols) chain<true>(Q, T, S, cq[right].data(), ct[right].data(), slot, H[right]);
		else chain<false>(Q, T, S, nullptr, nullptr, 0, H[right]);
	}
	finish(qstart, tstart, H[0], H[1], min_aln, out);
	if (want_cols && qstr) {
		const int first = out[7], size = out[5];
		if (size + 1 > cap) return -1;
		for (int c = 0; c < size; ++c) {                       // the merge of k_aln_pack (align.cu)
			const int m = first + c;
			if (m < H[0].cols) { qstr[c] = cq[0][(size_t)(H[0].cols - 1 - m)]; tstr[c] = ct[0][(size_t)(H[0].cols - 1 - m)]; }
			else { qstr[c] = cq[1][(size_t)(m - H[0].cols)]; tstr[c] = ct[1][(size_t)(m - H[0].cols)]; }
		}
		qstr[size] = 0; tstr[size] = 0;
	}
	return out[0];
}
