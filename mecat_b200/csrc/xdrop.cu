// mecat_b200/csrc/xdrop.cu -- the nanopore (`-x 1`) gapped extension on the GPU.
//
// CUDA backend of xdrop_core.cuh: XdropAligner::go (src/common/xdrop_gapalign.cpp:10-439) for a batch of candidates.
// A chain -- one (candidate, direction) -- is one thread's work: the X-drop row is a sequential scan (xdrop_core.cuh), so
// the parallelism is across the chains of a batch (2 x candidates of them); persistent warps take them 32 at a time from
// a list ordered by expected length, so that the lanes of a warp stay in step.  Every thread owns a worst-case-sized scratch block in global memory (score row + 4-bit trace-back of one
// block, 232 KB), so no chain can fail for want of memory.  Consumers: mecat2pw -j 1 -x 1 (string-free), mecat2ref -x 1
// and mecat_b200_align_batch policy 2 (columns into the slots of align.cu, merged and packed by its kernels).
#include "common.cuh"
#include "xdrop_core.cuh"

#include <algorithm>

namespace mb {

namespace {

constexpr int XD_THREADS = 128;
constexpr size_t XD_RING_BYTES = sizeof(mbx::RingCell) * mbx::RING * XD_THREADS;      // 64 KB per CTA: three CTAs per SM

struct TaskView { int32_t qread, qstrand, qstart, sread, sstart, swin_off, swin_len; };
__device__ __forceinline__ TaskView view_of(const AlignTask& t) { return {t.qread, t.qstrand, t.qstart, t.sread, t.sstart, t.swin_off, t.swin_len}; }
__device__ __forceinline__ TaskView view_of(const ExtendTask& t) { return {t.qread, t.qstrand, t.qstart, t.sread, t.sstart, 0, 0}; }

// one chain: the walks of align.cu for (task, direction), the block chain of xdrop_core.cuh, the slot's counters
template <bool COLS, class TaskT>
__device__ __forceinline__ void run_chain(const uint32_t* __restrict__ qfwd, const uint32_t* __restrict__ qrev, const int2* __restrict__ qoffsz, int qN,
                                          const uint32_t* __restrict__ sfwd, const uint32_t* __restrict__ srev, const int2* __restrict__ soffsz, int sN,
                                          const TaskT* __restrict__ tasks, unsigned long long item, AlnSlot* __restrict__ slots,
                                          char* __restrict__ colq, char* __restrict__ colt, const mbx::Scratch& S)
{
	using namespace mbx;
	const TaskView t = view_of(tasks[item >> 1]);
	const int right = (int)(item & 1);
	const int2 qo = qoffsz[t.qread];
	int2 so = soffsz[t.sread];
	if (t.swin_len > 0) { so.x += t.swin_off; so.y = t.swin_len; }    // subject window (mecat2ref)
	Seq Q, T;
	if (right) {
		if (!t.qstrand) { Q.arr = qfwd; Q.g0 = (uint32_t)(qo.x + t.qstart); Q.comp = 0; }
		else { Q.arr = qrev; Q.g0 = (uint32_t)(qN - qo.x - qo.y + t.qstart); Q.comp = 0xFFFFFFFFu; }
		Q.len = qo.y - t.qstart;
		T.arr = sfwd; T.g0 = (uint32_t)(so.x + t.sstart); T.comp = 0; T.len = so.y - t.sstart;
	} else {
		if (!t.qstrand) { Q.arr = qrev; Q.g0 = (uint32_t)(qN - qo.x - t.qstart); Q.comp = 0; }
		else { Q.arr = qfwd; Q.g0 = (uint32_t)(qo.x + qo.y - t.qstart); Q.comp = 0xFFFFFFFFu; }
		Q.len = t.qstart;
		T.arr = srev; T.g0 = (uint32_t)(sN - so.x - t.sstart); T.comp = 0; T.len = t.sstart;
	}
	AlnSlot slot = slots[item];
	Half H;
	chain<COLS>(Q, T, S, COLS ? colq + slot.off : nullptr, COLS ? colt + slot.off : nullptr, slot.cap, H);
	slot.cols = H.cols; slot.matches = H.matches; slot.qadv = H.qadv; slot.tadv = H.tadv;
	slot.overflow = H.overflow | (H.last << 1);
	slots[item] = slot;
}

template <bool COLS, class TaskT>
__global__ void __launch_bounds__(XD_THREADS)
k_xdrop(const uint32_t* __restrict__ qfwd, const uint32_t* __restrict__ qrev, const int2* __restrict__ qoffsz, int qN,
        const uint32_t* __restrict__ sfwd, const uint32_t* __restrict__ srev, const int2* __restrict__ soffsz, int sN,
        const TaskT* __restrict__ tasks, size_t ntasks, AlnSlot* __restrict__ slots, char* __restrict__ colq,
        char* __restrict__ colt, unsigned char* __restrict__ scratch, unsigned long long* __restrict__ work_counter,
        const uint32_t* __restrict__ order)
{
	using namespace mbx;
	const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned char* p = scratch + tid * SCRATCH_BYTES;
	Scratch S;
	S.sc = (Cell*)p; p += sizeof(Cell) * SC_CELLS;
	S.tb = (uint32_t*)p; p += 4 * (size_t)TB_WORDS;
	S.row_first = (int32_t*)p; p += 4 * (size_t)ROWS;
	S.row_word = (int32_t*)p;
	extern __shared__ RingCell ring_smem[];           // RING x blockDim entries: entry i of thread t at [i * blockDim + t]
	S.ring = ring_smem + threadIdx.x; S.ring_stride = XD_THREADS;      // (= blockDim.x; a constant folds into the addressing)
	// A warp takes 32 chains at a time from a list ordered by expected length (k_xd_class / k_xd_place), so that its lanes
	// start together, walk blocks of the same shape in step and finish at about the same time: a lane that asked for its
	// next chain on its own would never meet the others again (each lane a divergent path of its own, 1/32 of the warp).
	const int lane = (int)(threadIdx.x & 31u);
	for (;;) {
		unsigned long long base = 0;
		if (lane == 0) base = atomicAdd(work_counter, 32ull);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= 2 * ntasks) break;
		if (base + lane < 2 * ntasks)                      // (the lanes past the end idle through this last round)
			run_chain<COLS, TaskT>(qfwd, qrev, qoffsz, qN, sfwd, srev, soffsz, sN, tasks, order[base + lane], slots, colq, colt, S);
		__syncwarp();
	}
}

// Expected length class of a chain: the shorter of what is left of the two sequences in its direction, in steps of 256
// bases (64 classes, the last one open ended).
constexpr int XD_CLASSES = 64;
template <class TaskT>
__device__ __forceinline__ int chain_class(const TaskT& task, int right, const int2* __restrict__ qoffsz, const int2* __restrict__ soffsz)
{
	const TaskView t = view_of(task);
	const int ql = qoffsz[t.qread].y, sl = t.swin_len > 0 ? t.swin_len : soffsz[t.sread].y;
	const int len = right ? min(ql - t.qstart, sl - t.sstart) : min(t.qstart, t.sstart);
	return min(XD_CLASSES - 1, max(len, 0) >> 8);
}
template <class TaskT>
__global__ void k_xd_class(const TaskT* __restrict__ tasks, size_t nitems, const int2* __restrict__ qoffsz, const int2* __restrict__ soffsz,
                           unsigned int* __restrict__ hist)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < nitems) atomicAdd(hist + chain_class(tasks[i >> 1], (int)(i & 1), qoffsz, soffsz), 1u);
}
// hist -> first place of every class, longest class first (one thread: 64 entries)
__global__ void k_xd_scan(unsigned int* __restrict__ hist)
{
	unsigned int at = 0;
	for (int c = XD_CLASSES - 1; c >= 0; --c) { const unsigned int n = hist[c]; hist[c] = at; at += n; }
}
template <class TaskT>
__global__ void k_xd_place(const TaskT* __restrict__ tasks, size_t nitems, const int2* __restrict__ qoffsz, const int2* __restrict__ soffsz,
                           unsigned int* __restrict__ cursor, uint32_t* __restrict__ order)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < nitems) order[atomicAdd(cursor + chain_class(tasks[i >> 1], (int)(i & 1), qoffsz, soffsz), 1u)] = (uint32_t)i;
}

__global__ void k_xdrop_finalize(const ExtendTask* __restrict__ tasks, const AlnSlot* __restrict__ slots, size_t n, int min_aln,
                                 mecat_extend_result* __restrict__ out)
{
	const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	const ExtendTask t = tasks[i];
	const AlnSlot a = slots[2 * i], b = slots[2 * i + 1];
	mbx::Half L = {a.cols, a.matches, a.qadv, a.tadv, a.overflow >> 1, a.overflow & 1};
	mbx::Half R = {b.cols, b.matches, b.qadv, b.tadv, b.overflow >> 1, b.overflow & 1};
	int32_t o[8];
	mbx::finish(t.qstart, t.sstart, L, R, min_aln, o);
	mecat_extend_result r;
	r.ok = o[0]; r.qstart = o[1]; r.qend = o[2]; r.sstart = o[3]; r.send = o[4]; r.columns = o[5]; r.matches = o[6]; r.pad_ = 0;
	// XdropAligner::calc_ident, xdrop_gapalign.h:147-156: 100.0 * ident / n in IEEE double
	r.ident = r.columns ? __ddiv_rn(__dmul_rn(100.0, (double)r.matches), (double)r.columns) : 0.0;
	out[i] = r;
}

// persistent threads: as many as the chains can use and the device memory left allows (232 KB of scratch each)
int xdrop_threads(Ctx* c, size_t nchains, int* grid)
{
	int per_sm = 384;          // three CTAs of 128 threads: the shared-memory rings (64 KB per CTA) set the limit
	if (const char* e = getenv("MECAT_B200_XDROP_THREADS")) per_sm = std::max(XD_THREADS, atoi(e) / XD_THREADS * XD_THREADS);   // tuning hook
	size_t want = (size_t)c->sm_count * per_sm;
	want = std::min(want, (nchains + XD_THREADS - 1) / XD_THREADS * XD_THREADS);
	size_t free_b = 0, total_b = 0;
	if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
		size_t pooled = 0;
		for (auto& b : c->blocks) if (!b.used) pooled += b.bytes;
		const size_t room = (free_b + pooled) / 2;                      // leave half of what is left to the caller's arenas
		while (want > XD_THREADS && want * mbx::SCRATCH_BYTES > room) want = (want / 2 + XD_THREADS - 1) / XD_THREADS * XD_THREADS;
	}
	*grid = (int)(want / XD_THREADS);
	return 0;
}

template <bool COLS, class TaskT>
int xdrop_run(Ctx* c, const DVolume* q, const DVolume* s, const TaskT* d_tasks, size_t nb, AlnSlot* d_slots, char* d_colq, char* d_colt)
{
	if (!nb) return 0;
	int grid = 0;
	xdrop_threads(c, 2 * nb, &grid);
	unsigned char* d_scratch = nullptr;
	uint32_t* d_order = nullptr;
	unsigned int* d_hist = nullptr;
	if (2 * nb >= 0xFFFFFFFFull) MB_FAIL(c, "xdrop: %zu tasks in one batch", nb);
	auto body = [&]() -> int {
		MB_CUDA(c, c->dmalloc((void**)&d_scratch, (size_t)grid * XD_THREADS * mbx::SCRATCH_BYTES));
		MB_CUDA(c, c->alloc(&d_order, 2 * nb));
		MB_CUDA(c, c->alloc(&d_hist, (size_t)XD_CLASSES));
		MB_CUDA(c, cudaMemsetAsync(d_hist, 0, sizeof(unsigned int) * XD_CLASSES, c->stream));
		{
			KScope ks(c, MECAT_K_MERGE, 3);
			const unsigned g = (unsigned)((2 * nb + 255) / 256);
			k_xd_class<TaskT><<<g, 256, 0, c->stream>>>(d_tasks, 2 * nb, q->offsz, s->offsz, d_hist);
			k_xd_scan<<<1, 1, 0, c->stream>>>(d_hist);
			k_xd_place<TaskT><<<g, 256, 0, c->stream>>>(d_tasks, 2 * nb, q->offsz, s->offsz, d_hist, d_order);
		}
		MB_CUDA(c, cudaMemsetAsync(c->d_counters + 4, 0, 8, c->stream));
		MB_CUDA(c, cudaFuncSetAttribute(k_xdrop<COLS, TaskT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XD_RING_BYTES));
		{
			KScope ks(c, MECAT_K_EXTEND);
			k_xdrop<COLS, TaskT><<<grid, XD_THREADS, XD_RING_BYTES, c->stream>>>(q->fwd, q->rev, q->offsz, q->num_bases, s->fwd, s->rev, s->offsz, s->num_bases,
			                                                       d_tasks, nb, d_slots, d_colq, d_colt, d_scratch, c->d_counters + 4, d_order);
		}
		MB_CUDA(c, cudaGetLastError());
		return 0;
	};
	const int rc = body();
	c->dfree(d_scratch); c->dfree(d_order); c->dfree(d_hist);      // the pool only marks the blocks free; later work on the stream is ordered behind the kernel
	return rc;
}

}  // namespace

int xdrop_fill_slots(Ctx* c, const DVolume* q, const DVolume* s, const AlignTask* d_tasks, size_t nb, AlnSlot* d_slots,
                     char* d_colq, char* d_colt)
{
	return xdrop_run<true, AlignTask>(c, q, s, d_tasks, nb, d_slots, d_colq, d_colt);
}

int xdrop_extend(Ctx* c, const DVolume* q, const DVolume* s, const ExtendTask* d_tasks, size_t ntasks, int min_aln,
                 mecat_extend_result* d_res)
{
	if (!ntasks) return 0;
	AlnSlot* d_slots = nullptr;
	auto body = [&]() -> int {
		MB_CUDA(c, c->alloc(&d_slots, 2 * ntasks));
		MB_CUDA(c, cudaMemsetAsync(d_slots, 0, sizeof(AlnSlot) * 2 * ntasks, c->stream));
		if (xdrop_run<false, ExtendTask>(c, q, s, d_tasks, ntasks, d_slots, nullptr, nullptr)) return 1;
		{
			KScope ks(c, MECAT_K_FINAL);
			k_xdrop_finalize<<<(unsigned)((ntasks + 255) / 256), 256, 0, c->stream>>>(d_tasks, d_slots, ntasks, min_aln, d_res);
		}
		MB_CUDA(c, cudaGetLastError());
		return 0;
	};
	const int rc = body();
	c->dfree(d_slots);
	return rc;
}

}  // namespace mb
