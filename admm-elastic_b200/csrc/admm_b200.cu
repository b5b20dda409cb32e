// admm_b200.cu -- C-ABI (include/admm_b200.h) and host orchestration of the CUDA path.
// One handle = one GPU.  No CPU fallback: every entry point that computes needs the device.
#include "../../include/admm_b200.h"
#include "kernels.cuh"
#include "sptrsv.cuh"
#include "ldlt_blocks.hpp"
#include "sptrsv_blocks.cuh"
#include "uzawa.cuh"
#include "uzawa_blocks.cuh"
#include "mcgs_resident.cuh"
#include "mcgs_resident_f32.cuh"
#include "mcgs_owned_f32.cuh"
#include "mcgs_tiled_f32.cuh"
#include "dataflow_plan.hpp"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <memory>

using namespace admmb200;

namespace {

std::string g_create_error;

#define CK(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess) {                                                                   \
			throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_));          \
		}                                                                                          \
	} while (0)

template <typename T> struct DevBuf {
	T *p = nullptr;
	size_t n = 0;
	void alloc(size_t count) {
		release();
		n = count;
		if (count) { CK(cudaMalloc((void **)&p, count * sizeof(T))); }
	}
	void zero(cudaStream_t s) { if (n) CK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
	void upload(const T *h, size_t count, cudaStream_t s) {
		if (count > n) alloc(count);
		if (count) CK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
	}
	void upload(const std::vector<T> &h, cudaStream_t s) { alloc(h.size()); upload(h.data(), h.size(), s); }
	void release() { if (p) { cudaFree(p); p = nullptr; } n = 0; }
	~DevBuf() { release(); }
	DevBuf() {}
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
};

inline int pad32(int n) { return (n + 31) & ~31; }

struct TetBatchH {
	int n = 0, n_pad = 0, model = 0;
	double mu = 0, lambda = 0, kappa = 0, bulk = 0; // bulk: K of the prox penalty (<= 0: from mu, lambda)
	std::vector<int> idx;      // 4n
	std::vector<double> dminv; // 9n
	std::vector<double> w;     // n
	std::vector<int> row_off;  // n or empty
	size_t slot_base = 0;
	DevBuf<int4> d_idx;
	DevBuf<char> d_dminv, d_wdt2, d_u, d_z; // element precision, raw bytes
	DevBuf<char> d_q;                       // [4][n_pad] SVD warm start (quaternions), zero = no guess
	DevBuf<int> d_defer;                    // [2 + n]: counter, the queue of degenerate elements, consumer blocks done
};
struct TriBatchH {
	int n = 0, n_pad = 0;
	double limit_min = -100, limit_max = 100;
	std::vector<int> idx;      // 3n
	std::vector<double> rest;  // 4n
	std::vector<double> w;
	std::vector<int> row_off;
	size_t slot_base = 0;
	DevBuf<int4> d_idx;
	DevBuf<char> d_rest, d_wdt2, d_u, d_z;
};
struct PinsH {
	int n = 0;
	std::vector<int> idx;
	std::vector<double> pos, w;
	std::vector<unsigned char> active;
	std::vector<int> row_off;
	DevBuf<int> d_idx;
	DevBuf<double> d_pos, d_wdt2, d_u, d_z;
	DevBuf<unsigned char> d_active;
	DevBuf<double4> d_f;
};

} // namespace

struct admm_b200_solver {
	int device = 0;
	int n_sms = 0;
	int gs_parts = 0; // parts (= CTAs) of the resident Gauss-Seidel on this GPU: one per SM unless admm_b200_set_gs_parts asked for fewer
	cudaStream_t own_stream = nullptr, stream = nullptr;
	std::string err;
	long long launches = 0;

	// nodes
	int n_nodes = 0;
	std::vector<double> h_m; // 3n
	DevBuf<double4> x, v, cx, mxbar, b, m;
	DevBuf<double> stage3, stage3b; // 3n staging for AoS host copies

	// elements
	std::vector<TetBatchH *> tets;
	std::vector<TriBatchH *> tris;
	PinsH pins;
	size_t n_slots = 0;
	DevBuf<char> f; // Vec4<E> per slot
	DevBuf<int> inc_off, inc_slot;

	// system
	int sys_n = 0;
	std::vector<int> L_rowptr, L_cols;
	std::vector<double> L_vals;
	std::vector<int> color_off, color_nodes;
	int n_colors = 0;
	// gs pins / obstacles
	std::vector<int> gs_pin_idx;
	std::vector<double> gs_pin_pos;
	DevBuf<int> d_pin_slot;
	DevBuf<double> d_pin_pos;
	std::vector<Obstacle> obstacles;
	// Solver::ext_forces (src/Solver.hpp:71): wind forces, applied in order at the top of every step
	struct Wind {
		std::vector<int> tris; double dir[3]; bool uploaded = false; int n_touched = 0;
		DevBuf<int> d_tris, d_nodes, d_inc_ptr, d_inc_tri; DevBuf<double4> d_kick;
	};
	std::vector<std::unique_ptr<Wind>> winds;
	DevBuf<Obstacle> d_obstacles;

	// mcgs device
	int gs_lanes = 4;
	DevBuf<int> gs_color_first, gs_slice_ptr, gs_slice_node, gs_ell_col;
	DevBuf<double> gs_ell_val, gs_diag, gs_resid;
	DevBuf<unsigned int> barrier;
	DevBuf<int> gs_iters_done, iter_log;
	// UzawaCG with passive collisions (uzawa.cuh)
	DevBuf<int> uz_hv, uz_ctl; DevBuf<double> uz_hn, uz_hc, uz_y, uz_r, uz_d, uz_q3, uz_scal; DevBuf<double4> uz_q1, uz_q2;
	int uz_max_iters = 20; double uz_tol = 1e-10; // src/UzawaCG.hpp:45-46
	std::vector<int> surface_inds;   // Solver::surface_inds: the vertices Collider::detect tests, in that order (empty: all)
	double constraint_w = 1.0;       // ConstraintSet::constraint_w of the UzawaCG path (src/Solver.cpp:239,245)
	DevBuf<int> uz_cand, uz_cta_cnt; DevBuf<double> uz_red;
	int gs_grid = 0;
	size_t gs_nnz = 0, gs_ell_entries = 0;
	// mcgs, shared-memory-resident variant (mcgs_resident.cuh)
	bool gs_resident = false;
	int gs_res_lanes = 1;
	int gs_owned_threads = 0;  // > 0: mcgs_owned_f32_kernel with that many threads (round 1's fp32 solve; obstacles, unequal masses)
	bool gs_tiled = false;     // mcgs_tiled_f32_kernel (the production fp32 solve): short tasks of 8 nodes x 4 lanes
	DevBuf<float> res_val_scaled; DevBuf<uint4> res_dest4;
	size_t gs_res_smem = 0;
	DevBuf<PartDesc> res_parts;
	DevBuf<uint16_t> res_col;
	DevBuf<char> res_val;
	DevBuf<int> res_gid, res_slice_row, res_color_slice, res_nbr, res_halo_color;
	DevBuf<float4> res_nodebuf;
	DevBuf<uint2> res_dglob;
	DevBuf<int> res_dest_off; DevBuf<unsigned int> res_dest_slot; int res_total_slots = 0; // mailboxes (plan_mailboxes)
	unsigned int gs_solve_seq = 0;
	DevBuf<double> res_val64;
	DevBuf<unsigned long long> res_prof; // ADMM_B200_GS_PROF=1: per-part cycle counters of the last solve
	DevBuf<unsigned int> res_sync; // part_epoch [8 * n_sms] | sweep_flag [iters] | sweep_arrive [iters]
	// everything the resident solve needs zeroed per launch, in one allocation (one memset per solve):
	// residual slots (doubles) | grid barrier counter (padded) | res_sync layout
	DevBuf<double> res_scratch; size_t res_scratch_resid_n = 0;
	// kernel-only timing (admm_b200_kernel_times)
	bool fine_on = false; std::vector<cudaEvent_t> fine_pool; size_t fine_used = 0; std::vector<int> fine_kind;
	double kernel_ms[3] = {0, 0, 0}; long long kernel_n[3] = {0, 0, 0};
	struct TimerBlockT { size_t ev0; int iters; size_t fine0, fine1; int log0; };
	std::vector<TimerBlockT> tblocks; size_t ev_next = 0; int log_next = 0, timer_steps = 0; bool deferred_timers = false; int deferred_stride = 1; long long deferred_count = 0;
	admm_b200_runtime pend = {0, 0, 0, 0, 0, 0}; double pend_kernel_ms[3] = {0, 0, 0}; long long pend_kernel_n[3] = {0, 0, 0}; int pend_steps = 0;
	DevBuf<short> res_slice_node;
	std::vector<double> h_x0; // rest positions (partitioning)
	std::string gs_info;
	// multi-GPU: one handle per rank, peers' buffers mapped through CUDA IPC
	int rank = 0, world = 1;
	DevBuf<unsigned int> mg_dest_mask, mg_flags; // flags: [ADMMB200_MAX_RANKS * 8], slot 8*q = epoch of rank q's last barrier
	uint2 *peer_dglob[ADMMB200_MAX_RANKS] = {nullptr};
	double4 *peer_x[ADMMB200_MAX_RANKS] = {nullptr};
	unsigned int *peer_flags[ADMMB200_MAX_RANKS] = {nullptr};
	std::vector<void *> ipc_opened;
	// multi-GPU step_host: only the nodes this rank touches travel -- owned + ghost nodes up, owned nodes down
	std::vector<int> mg_local, mg_owned;     // node ids, ascending
	std::vector<int> mg_runs;                // mg_local as runs of consecutive ids: {first id, length, offset in mg_local} x n_runs
	DevBuf<int> d_mg_local, d_mg_owned;
	double *mg_pinned = nullptr;             // pinned host staging: [x local | v local]
	bool mg_ready = false;
	unsigned int mg_epoch = 0;

	// ldlt
	bool have_ldlt = false;
	int ld_n = 0;
	std::vector<int> ld_perm, ld_Lp, ld_Li;
	std::vector<double> ld_Lx, ld_D;
	DevBuf<int> d_ld_perm, d_fwd_level_ptr, d_fwd_rows, d_fwd_rowptr, d_fwd_cols, d_bwd_level_ptr, d_bwd_rows, d_bwd_rowptr, d_bwd_cols;
	DevBuf<double> d_fwd_vals, d_bwd_vals, d_ld_D;
	DevBuf<double4> d_ld_y;
	int ld_levels_fwd = 0, ld_levels_bwd = 0, ld_grid = 0, ld_lanes = 4;
	// block (supernodal) solve, sptrsv_blocks.cuh -- the default; the level-scheduled kernel above stays as ADMM_B200_LDLT_KERNEL=levels
	bool ld_blocks = false;
	int lb_levels = 0, lb_cut = 0;
	DevBuf<int> lb_blk_of, lb_blk_c0, lb_lev_ptr, lb_rows, lb_lanes, lb_f_rowptr, lb_f_cols, lb_b_colptr, lb_b_rows, lb_seg_ptr, lb_seg_begin, lb_seg_end, lb_seg_level;
	DevBuf<long long> lb_inv_off;
	DevBuf<double> lb_inv, lb_invT, lb_f_vals, lb_b_vals;
	DevBuf<double4> lb_t;
	DevBuf<int4> lb_desc_fg, lb_desc_bg, lb_desc_d; DevBuf<long long> lb_desc_off;
	DevBuf<unsigned long long> lb_prof; // ADMM_B200_LDLT_PROF=1: phase stamps of the last block solve (debug_get "ldlt_prof")

	// settings
	bool finalized = false;
	double dt = 1.0 / 24.0;
	int linsolver = 0, gs_iters = 30, precision = 0;
	double gs_omega = 1.9, gs_tol = 1e-10;
	bool store_z = false;
	// resident blocks/SM the fp32 tet kernel is compiled for (ADMM_B200_TET_MINBLOCKS = 5, 6, 8).  Measured local phase on
	// the 1M-tet beam: 5 (96 regs) 61.6 us, 6 (80) 57.5, 8 (64 regs, 44 B spilled) 55.4, 10 (48) 63.4, 12 (40) 78.0: the kernel
	// hides its index -> gather latency with resident warps; two elements per thread with hand-hoisted loads lost (58.9).
	int tet_minblocks = 8;

	// events
	std::vector<cudaEvent_t> events;

	~admm_b200_solver() {
		for (void *p : ipc_opened) cudaIpcCloseMemHandle(p);
		if (mg_pinned) cudaFreeHost(mg_pinned);
		for (auto t : tets) delete t;
		for (auto t : tris) delete t;
		for (auto e : events) cudaEventDestroy(e);
		if (own_stream) cudaStreamDestroy(own_stream);
	}
};

namespace {

typedef admm_b200_solver S;
typedef admm_b200_solver::TimerBlockT TimerBlock;
void fine_begin(S *s, int kind);
void fine_end(S *s);

template <typename F> int guard(S *s, F fn)
{
	try {
		if (s) CK(cudaSetDevice(s->device));
		fn();
		return 0;
	} catch (std::exception &e) {
		if (s) s->err = e.what(); else g_create_error = e.what();
		return 1;
	}
}

inline void require(bool c, const char *msg) { if (!c) throw std::runtime_error(msg); }

template <typename E> void upload_soa(DevBuf<char> &dst, const std::vector<double> &src, int n, int n_pad, int k, cudaStream_t st, double scale_unused = 1.0)
{
	// src is AoS [n][k] doubles -> dst SoA [k][n_pad] of E
	std::vector<E> tmp((size_t)k * n_pad, E(0));
	for (int e = 0; e < n; ++e)
		for (int j = 0; j < k; ++j) tmp[(size_t)j * n_pad + e] = E(src[(size_t)e * k + j]);
	dst.alloc(tmp.size() * sizeof(E));
	CK(cudaMemcpyAsync(dst.p, tmp.data(), tmp.size() * sizeof(E), cudaMemcpyHostToDevice, st));
	CK(cudaStreamSynchronize(st)); // tmp dies at scope exit
}

// ------------------------------------------------------------------------------------------
// local-step launches
// ------------------------------------------------------------------------------------------
template <typename E, int MODEL> void launch_tet_model(S *s, TetBatchH *t)
{
	TetBatch<E> tb;
	tb.n = t->n; tb.n_pad = t->n_pad;
	tb.idx = t->d_idx.p;
	tb.dminv = (const E *)t->d_dminv.p;
	tb.wdt2 = (const E *)t->d_wdt2.p;
	tb.u = (E *)t->d_u.p;
	tb.z = (E *)t->d_z.p;
	tb.q = (E *)t->d_q.p;
	tb.f = (typename Vec4<E>::type *)s->f.p + t->slot_base;
	tb.mat = Material<E>::make(t->mu, t->lambda, t->kappa, t->bulk);
	tb.defer_count = t->d_defer.p; tb.defer_list = t->d_defer.p + 1; tb.defer_done = t->d_defer.p + 1 + t->n;
	const int threads = 128;
	int blocks = (t->n + threads - 1) / threads;
	const bool sz = s->store_z && t->d_z.p;
	fine_begin(s, 0);
	if (sz) tet_local_kernel<E, MODEL, true, 4><<<blocks, threads, 0, s->stream>>>(tb, s->cx.p);
	else if (sizeof(E) == 8) tet_local_kernel<E, MODEL, false, 4><<<blocks, threads, 0, s->stream>>>(tb, s->cx.p);
	else if (s->tet_minblocks >= 10) tet_local_kernel<E, MODEL, false, 10><<<blocks, threads, 0, s->stream>>>(tb, s->cx.p);
	else if (s->tet_minblocks >= 8) tet_local_kernel<E, MODEL, false, 8><<<blocks, threads, 0, s->stream>>>(tb, s->cx.p);
	else if (s->tet_minblocks >= 6) tet_local_kernel<E, MODEL, false, 6><<<blocks, threads, 0, s->stream>>>(tb, s->cx.p);
	else tet_local_kernel<E, MODEL, false, 5><<<blocks, threads, 0, s->stream>>>(tb, s->cx.p);
	CK(cudaGetLastError());
	fine_end(s);
	s->launches++;
	if (MODEL != TET_LINEAR) {
		// degenerate elements queued by the kernel above (usually none)
		const int dblocks = std::min(blocks, 2 * s->n_sms);
		if (sz) tet_local_deferred_kernel<E, MODEL, true><<<dblocks, threads, 0, s->stream>>>(tb, s->cx.p);
		else tet_local_deferred_kernel<E, MODEL, false><<<dblocks, threads, 0, s->stream>>>(tb, s->cx.p);
		CK(cudaGetLastError());
		s->launches++;
	}
}

template <typename E> void launch_tet(S *s, TetBatchH *t)
{
	switch (t->model) {
	case TET_LINEAR: launch_tet_model<E, TET_LINEAR>(s, t); break;
	case TET_NEOHOOKEAN: launch_tet_model<E, TET_NEOHOOKEAN>(s, t); break;
	case TET_STVK: launch_tet_model<E, TET_STVK>(s, t); break;
	case TET_SPLINE_NH: launch_tet_model<E, TET_SPLINE_NH>(s, t); break;
	case TET_SPLINE_STVK: launch_tet_model<E, TET_SPLINE_STVK>(s, t); break;
	case TET_SPLINE_COROT: launch_tet_model<E, TET_SPLINE_COROT>(s, t); break;
	default: throw std::runtime_error("unknown tet model");
	}
}

template <typename E> void launch_tri(S *s, TriBatchH *t)
{
	TriBatch<E> tb;
	tb.n = t->n; tb.n_pad = t->n_pad;
	tb.idx = t->d_idx.p;
	tb.rest = (const E *)t->d_rest.p;
	tb.wdt2 = (const E *)t->d_wdt2.p;
	tb.u = (E *)t->d_u.p;
	tb.z = (E *)t->d_z.p;
	tb.f = (typename Vec4<E>::type *)s->f.p + t->slot_base;
	tb.limit_min = E(t->limit_min); tb.limit_max = E(t->limit_max);
	const int threads = 128;
	int blocks = (t->n + threads - 1) / threads;
	fine_begin(s, 0);
	if (s->store_z && t->d_z.p) tri_local_kernel<E, true><<<blocks, threads, 0, s->stream>>>(tb, s->cx.p);
	else tri_local_kernel<E, false><<<blocks, threads, 0, s->stream>>>(tb, s->cx.p);
	CK(cudaGetLastError());
	fine_end(s);
	s->launches++;
}

void launch_pins(S *s)
{
	if (!s->pins.n) return;
	PinBatch pb;
	pb.n = s->pins.n; pb.idx = s->pins.d_idx.p; pb.pos = s->pins.d_pos.p; pb.active = s->pins.d_active.p;
	pb.wdt2 = s->pins.d_wdt2.p; pb.u = s->pins.d_u.p; pb.z = s->pins.d_z.p; pb.f = s->pins.d_f.p;
	pin_local_kernel<<<(pb.n + 127) / 128, 128, 0, s->stream>>>(pb, s->cx.p);
	CK(cudaGetLastError());
	s->launches++;
}

void launch_local(S *s)
{
	for (auto t : s->tets) { if (s->precision == ADMM_B200_FP64) launch_tet<double>(s, t); else launch_tet<float>(s, t); }
	for (auto t : s->tris) { if (s->precision == ADMM_B200_FP64) launch_tri<double>(s, t); else launch_tri<float>(s, t); }
	launch_pins(s);
}

void launch_assemble(S *s)
{
	int n = s->n_nodes;
	fine_begin(s, 1);
	if (s->precision == ADMM_B200_FP64)
		assemble_kernel<double><<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->inc_off.p, s->inc_slot.p, (const double4 *)s->f.p, s->pins.d_f.p, s->mxbar.p, s->b.p);
	else
		assemble_kernel<float><<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->inc_off.p, s->inc_slot.p, (const float4 *)s->f.p, s->pins.d_f.p, s->mxbar.p, s->b.p);
	CK(cudaGetLastError());
	fine_end(s);
	s->launches++;
}

// ------------------------------------------------------------------------------------------
// global solve launches
// ------------------------------------------------------------------------------------------
template <int T> void mcgs_launch_T(S *s, McgsParams &P)
{
	void *args[] = {&P};
	CK(cudaLaunchCooperativeKernel((void *)mcgs_kernel<T>, dim3(s->gs_grid), dim3(ADMMB200_MCGS_THREADS), args, 0, s->stream));
}
template <int T> int mcgs_occupancy()
{
	int nb = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mcgs_kernel<T>, ADMMB200_MCGS_THREADS, 0));
	return nb;
}

void fill_mcgs_params(S *s, McgsParams &P, bool zero_scratch = true);

// precision FP64: positions in shared memory as doubles (mcgs_resident_kernel<double>);
// precision FP32: fp32 sweeps on the increment around the fp64 anchor (mcgs_resident_f32_kernel)
const void *resident_kernel_ptr(bool fp64, int lanes, bool prof)
{
	if (fp64) {
		if (lanes == 1) return (const void *)mcgs_resident_kernel<double, 1>;
		if (lanes == 2) return (const void *)mcgs_resident_kernel<double, 2>;
		return (const void *)mcgs_resident_kernel<double, 4>;
	}
	if (prof) {
		if (lanes == 1) return (const void *)mcgs_resident_f32_kernel<1, true>;
		if (lanes == 2) return (const void *)mcgs_resident_f32_kernel<2, true>;
		return (const void *)mcgs_resident_f32_kernel<4, true>;
	}
	if (lanes == 1) return (const void *)mcgs_resident_f32_kernel<1, false>;
	if (lanes == 2) return (const void *)mcgs_resident_f32_kernel<2, false>;
	return (const void *)mcgs_resident_f32_kernel<4, false>;
}

// mcgs_owned_f32_kernel<threads, slices per warp, obstacles, profiling>: 512 x 4 and 768 x 3 slices
const void *owned_kernel_ptr(int threads, bool obst, bool prof)
{
	if (threads == 512) {
		if (prof) return obst ? (const void *)mcgs_owned_f32_kernel<512, 4, true, true> : (const void *)mcgs_owned_f32_kernel<512, 4, false, true>;
		return obst ? (const void *)mcgs_owned_f32_kernel<512, 4, true, false> : (const void *)mcgs_owned_f32_kernel<512, 4, false, false>;
	}
	if (prof) return obst ? (const void *)mcgs_owned_f32_kernel<768, 3, true, true> : (const void *)mcgs_owned_f32_kernel<768, 3, false, true>;
	return obst ? (const void *)mcgs_owned_f32_kernel<768, 3, true, false> : (const void *)mcgs_owned_f32_kernel<768, 3, false, false>;
}
inline int owned_capacity(int threads) { return threads == 512 ? 4 * 16 : 3 * 24; }

void launch_mcgs_resident(S *s)
{
	const bool fp64 = s->precision == ADMM_B200_FP64;
	McgsResParams R;
	McgsRes32Params R32;
	McgsParams &B = fp64 ? R.base : R32.base;
	fill_mcgs_params(s, B, false);
	// scratch: [resid doubles][barrier: 2 uints][part_epoch 8 n_sms | sweep_flag iters | sweep_arrive iters]
	CK(cudaMemsetAsync(s->res_scratch.p, 0, s->res_scratch.n * sizeof(double), s->stream));
	B.resid = s->res_scratch.p; B.resid_lb = s->res_scratch.p + (s->gs_iters + 2);
	unsigned int *scr_u = (unsigned int *)(s->res_scratch.p + s->res_scratch_resid_n);
	B.barrier = scr_u;
	unsigned int *part_epoch = scr_u + 2, *sweep_flag = part_epoch + 8 * (size_t)s->gs_parts, *sweep_arrive = sweep_flag + s->gs_iters;
	void *args[1];
	if (fp64) {
		R.parts = s->res_parts.p; R.col = s->res_col.p; R.val = s->res_val.p; R.gid = s->res_gid.p;
		R.slice_row = s->res_slice_row.p; R.color_slice = s->res_color_slice.p; R.slice_node = s->res_slice_node.p; R.nbr = s->res_nbr.p;
		R.part_epoch = part_epoch; R.sweep_flag = sweep_flag; R.sweep_arrive = sweep_arrive; R.prof = s->res_prof.p;
		args[0] = &R;
	} else {
		R32.parts = s->res_parts.p; R32.col = s->res_col.p; R32.val = (const float *)s->res_val.p; R32.gid = s->res_gid.p;
		R32.slice_row = s->res_slice_row.p; R32.color_slice = s->res_color_slice.p; R32.slice_node = s->res_slice_node.p; R32.nbr = s->res_nbr.p;
		R32.halo_color = s->res_halo_color.p;
		R32.part_epoch = part_epoch; R32.sweep_flag = sweep_flag; R32.sweep_arrive = sweep_arrive; R32.prof = s->res_prof.p;
		R32.dglob = s->res_dglob.p; R32.nodebuf = s->res_nodebuf.p; R32.val64 = s->res_val64.p;
		R32.dest_off = s->res_dest_off.p; R32.dest_slot = s->res_dest_slot.p; R32.total_slots = s->res_total_slots;
		s->gs_solve_seq = (s->gs_solve_seq + 1) & 0xFFFFFu;
		if (s->gs_solve_seq == 0) s->gs_solve_seq = 1;
		R32.tag_base = s->gs_solve_seq << 12;
		R32.n_nodes_total = s->n_nodes;
		R32.part0 = s->rank * s->gs_parts; R32.world = s->world; R32.rank = s->rank;
		{ static const char *dbg = getenv("ADMM_B200_GS_DBG"); R32.dbg = dbg ? atoi(dbg) : 0; }
		R32.dest_mask = s->world > 1 ? s->mg_dest_mask.p : nullptr;
		for (int q = 0; q < ADMMB200_MAX_RANKS; ++q) { R32.peer_dglob[q] = s->peer_dglob[q]; R32.peer_x[q] = s->peer_x[q]; }
		if (s->world > 1) { require(s->mg_ready, "multi-GPU solver used before admm_b200_mgpu_ready"); R32.base.tol2 = 0.0; }
		args[0] = &R32;
	}
	fine_begin(s, 2);
	if (!fp64 && s->gs_tiled) {
		McgsTiledExtra X; X.val_scaled = s->res_val_scaled.p; X.dest4 = s->res_dest4.p;
		void *targs[2] = {&R32, &X};
		const void *kern = s->res_prof.p ? (const void *)mcgs_tiled_f32_kernel<4, true> : (const void *)mcgs_tiled_f32_kernel<4, false>;
		CK(cudaLaunchCooperativeKernel(kern, dim3(s->gs_parts), dim3(ADMMB200_TILED_THREADS), targs, s->gs_res_smem, s->stream));
	} else if (!fp64 && s->gs_owned_threads > 0) {
		const void *kern = owned_kernel_ptr(s->gs_owned_threads, !s->obstacles.empty(), s->res_prof.p != nullptr);
		CK(cudaLaunchCooperativeKernel(kern, dim3(s->gs_parts), dim3(s->gs_owned_threads), args, s->gs_res_smem, s->stream));
	} else
		CK(cudaLaunchCooperativeKernel(resident_kernel_ptr(fp64, s->gs_res_lanes, s->res_prof.p != nullptr), dim3(s->gs_parts), dim3(ADMMB200_RES_THREADS), args, s->gs_res_smem, s->stream));
	fine_end(s);
	s->launches++;
}

void fill_mcgs_params(S *s, McgsParams &P, bool zero_scratch)
{
	P.n_nodes = s->n_nodes; P.n_colors = s->n_colors; P.iters = s->gs_iters; P.omega = s->gs_omega;
	P.tol2 = s->gs_tol > 0 ? s->gs_tol * s->gs_tol : 0.0;
	P.color_first_slice = s->gs_color_first.p; P.slice_ptr = s->gs_slice_ptr.p; P.slice_node = s->gs_slice_node.p;
	P.ell_col = s->gs_ell_col.p; P.ell_val = s->gs_ell_val.p; P.diag = s->gs_diag.p;
	P.pin_slot = s->d_pin_slot.p; P.pin_pos = s->d_pin_pos.p; P.has_pins = s->gs_pin_idx.empty() ? 0 : 1;
	P.n_obstacles = (int)s->obstacles.size();
	P.obs = s->d_obstacles.p;
	P.x = s->cx.p; P.b = s->b.p; P.barrier = s->barrier.p; P.resid = s->gs_resid.p; P.resid_lb = s->gs_resid.p + (s->gs_iters + 2); P.iters_done = s->gs_iters_done.p;
	if (!zero_scratch) return; // the resident solve zeroes its own scratch block in one go
	CK(cudaMemsetAsync(s->barrier.p, 0, sizeof(unsigned int), s->stream));
	if (P.tol2 > 0) CK(cudaMemsetAsync(s->gs_resid.p, 0, s->gs_resid.n * sizeof(double), s->stream));
}

void launch_mcgs(S *s)
{
	if (s->gs_resident) { launch_mcgs_resident(s); return; }
	McgsParams P;
	fill_mcgs_params(s, P);
	fine_begin(s, 2);
	switch (s->gs_lanes) {
	case 1: mcgs_launch_T<1>(s, P); break;
	case 2: mcgs_launch_T<2>(s, P); break;
	case 8: mcgs_launch_T<8>(s, P); break;
	default: mcgs_launch_T<4>(s, P); break;
	}
	fine_end(s);
	s->launches++;
}

template <int T> void ldlt_launch_T(S *s, LdltParams &P)
{
	void *args[] = {&P};
	CK(cudaLaunchCooperativeKernel((void *)ldlt_solve_kernel<T>, dim3(s->ld_grid), dim3(512), args, 0, s->stream));
}

void fill_ldlt_blocks(S *s, LdltBlkParams &B)
{
	B.n = s->ld_n; B.n_levels = s->lb_levels; B.cut = s->lb_cut;
	B.perm = s->d_ld_perm.p; B.blk_of = s->lb_blk_of.p; B.blk_c0 = s->lb_blk_c0.p; B.inv_off = s->lb_inv_off.p; B.inv = s->lb_inv.p; B.invT = s->lb_invT.p;
	B.lev_ptr = s->lb_lev_ptr.p; B.rows = s->lb_rows.p; B.lanes = s->lb_lanes.p;
	B.f_rowptr = s->lb_f_rowptr.p; B.f_cols = s->lb_f_cols.p; B.f_vals = s->lb_f_vals.p;
	B.b_colptr = s->lb_b_colptr.p; B.b_rows = s->lb_b_rows.p; B.b_vals = s->lb_b_vals.p;
	B.seg_ptr = s->lb_seg_ptr.p; B.seg_begin = s->lb_seg_begin.p; B.seg_end = s->lb_seg_end.p; B.seg_level = s->lb_seg_level.p;
	B.D = s->d_ld_D.p; B.t = s->lb_t.p; B.y = s->d_ld_y.p; B.b = nullptr; B.x = nullptr; B.barrier = s->barrier.p; B.active = nullptr;
	B.prof = s->lb_prof.p;
	B.desc_fg = s->lb_desc_fg.p; B.desc_bg = s->lb_desc_bg.p; B.desc_d = s->lb_desc_d.p; B.desc_off = s->lb_desc_off.p;
}

void launch_ldlt(S *s, const double4 *rhs = nullptr, double4 *out = nullptr, const int *active = nullptr)
{
	LdltParams P;
	P.n = s->ld_n; P.n_levels_fwd = s->ld_levels_fwd; P.n_levels_bwd = s->ld_levels_bwd;
	P.perm = s->d_ld_perm.p;
	P.fwd_level_ptr = s->d_fwd_level_ptr.p; P.fwd_rows = s->d_fwd_rows.p; P.fwd_rowptr = s->d_fwd_rowptr.p; P.fwd_cols = s->d_fwd_cols.p; P.fwd_vals = s->d_fwd_vals.p;
	P.bwd_level_ptr = s->d_bwd_level_ptr.p; P.bwd_rows = s->d_bwd_rows.p; P.bwd_rowptr = s->d_bwd_rowptr.p; P.bwd_cols = s->d_bwd_cols.p; P.bwd_vals = s->d_bwd_vals.p;
	P.dinv_unused = nullptr; P.D = s->d_ld_D.p; P.y = s->d_ld_y.p; P.b = rhs ? rhs : s->b.p; P.x = out ? out : s->cx.p; P.barrier = s->barrier.p; P.active = active;
	CK(cudaMemsetAsync(s->barrier.p, 0, sizeof(unsigned int), s->stream));
	if (s->ld_blocks) {
		LdltBlkParams B;
		fill_ldlt_blocks(s, B);
		B.b = P.b; B.x = P.x; B.active = active;
		void *args[] = {&B};
		fine_begin(s, 2);
		CK(cudaLaunchCooperativeKernel((void *)ldlt_blocks_kernel, dim3(s->ld_grid), dim3(1024), args, 0, s->stream));
		fine_end(s);
		s->launches++;
		return;
	}
	fine_begin(s, 2);
	switch (s->ld_lanes) {
	case 1: ldlt_launch_T<1>(s, P); break;
	case 2: ldlt_launch_T<2>(s, P); break;
	case 8: ldlt_launch_T<8>(s, P); break;
	default: ldlt_launch_T<4>(s, P); break;
	}
	fine_end(s);
	s->launches++;
}

// Cross-GPU barrier (one tiny kernel): every rank stores its epoch into every peer's flag array and
// waits until all peers' epochs have arrived in its own.  Runs after the solve so that the ghost
// positions the peers pushed at the end of their solve are complete before the next local step.
struct MgBarrierParams { int world, rank; unsigned int epoch; unsigned int *mine; unsigned int *peer[ADMMB200_MAX_RANKS]; };
__global__ void mgpu_barrier_kernel(MgBarrierParams B)
{
	const int q = threadIdx.x;
	if (q >= B.world || q == B.rank) return;
	__threadfence_system();
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(B.peer[q] + 8 * B.rank), "r"(B.epoch) : "memory");
	unsigned int seen;
	do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(B.mine + 8 * q) : "memory"); } while (seen < B.epoch);
}

void launch_mgpu_barrier(S *s)
{
	if (s->world <= 1) return;
	require(s->mg_ready, "multi-GPU solver used before admm_b200_mgpu_ready");
	MgBarrierParams B;
	B.world = s->world; B.rank = s->rank; B.epoch = ++s->mg_epoch; B.mine = s->mg_flags.p;
	for (int q = 0; q < ADMMB200_MAX_RANKS; ++q) B.peer[q] = s->peer_flags[q];
	mgpu_barrier_kernel<<<1, 32, 0, s->stream>>>(B);
	CK(cudaGetLastError());
	s->launches++;
}

// UzawaCG::solve with passive collisions (uzawa.cuh): a fixed-length sequence of launches, ended early on the device
void launch_uzawa(S *s)
{
	const int n = s->n_nodes, nb = (n + 255) / 256;
	UzParams U;
	U.n = n; U.n_obstacles = (int)s->obstacles.size(); U.obs = s->d_obstacles.p;
	U.hv = s->uz_hv.p; U.hn = s->uz_hn.p; U.hc = s->uz_hc.p; U.y = s->uz_y.p; U.r = s->uz_r.p; U.d = s->uz_d.p; U.q3 = s->uz_q3.p;
	U.ctl = s->uz_ctl.p; U.scal = s->uz_scal.p; U.tol2 = s->uz_tol * s->uz_tol;
	if (s->ld_blocks) {
		// one persistent cooperative launch: detection, the conjugate-gradient loop and every A^-1 inside (uzawa_blocks.cuh)
		UzBlkParams Z;
		fill_ldlt_blocks(s, Z.L);
		Z.U = U;
		Z.cand = s->surface_inds.empty() ? nullptr : s->uz_cand.p; Z.n_cand = (int)s->surface_inds.size();
		Z.ck = std::sqrt(std::max(0.0, s->constraint_w)); Z.max_iters = s->uz_max_iters;
		Z.x = s->cx.p; Z.b = s->b.p; Z.q1 = s->uz_q1.p; Z.q2 = s->uz_q2.p; Z.cta_cnt = s->uz_cta_cnt.p; Z.red = s->uz_red.p; Z.iters_done = s->gs_iters_done.p;
		CK(cudaMemsetAsync(s->barrier.p, 0, sizeof(unsigned int), s->stream));
		void *args[] = {&Z};
		fine_begin(s, 2);
		CK(cudaLaunchCooperativeKernel((void *)uzawa_blocks_kernel, dim3(s->ld_grid), dim3(1024), args, 0, s->stream));
		fine_end(s);
		s->launches++;
		return;
	}
	require(s->surface_inds.empty() && s->constraint_w == 1.0, "surface_inds / constraint_w need the block UzawaCG kernel (unset ADMM_B200_LDLT_KERNEL=levels)");
	const int *active = s->uz_ctl.p + 2;
	// hits at the current iterate (Solver::step, src/Solver.cpp:87-90), then x = A^-1 (b - C^T y) (:83-84)
	uz_detect_kernel<<<1, 1024, 0, s->stream>>>(U, s->cx.p);
	uz_copy_kernel<<<nb, 256, 0, s->stream>>>(n, s->b.p, s->uz_q1.p);
	uz_scatter_kernel<<<nb, 256, 0, s->stream>>>(U, s->uz_y.p, -1.0, s->uz_q1.p, 0);
	CK(cudaGetLastError());
	s->launches += 3;
	launch_ldlt(s, s->uz_q1.p, s->cx.p, nullptr);
	uz_init_kernel<<<nb, 256, 0, s->stream>>>(U, s->cx.p);
	s->launches++;
	for (int it = 0; it < s->uz_max_iters; ++it) {
		// q2 = A^-1 C^T d; alpha, y, r, beta, d on the rows; x -= alpha q2   (:93-118)
		uz_copy_kernel<<<nb, 256, 0, s->stream>>>(n, nullptr, s->uz_q1.p);
		uz_scatter_kernel<<<nb, 256, 0, s->stream>>>(U, s->uz_d.p, 1.0, s->uz_q1.p, 1);
		launch_ldlt(s, s->uz_q1.p, s->uz_q2.p, active);
		uz_step_kernel<<<1, 1024, 0, s->stream>>>(U, s->uz_q2.p);
		uz_axpy_kernel<<<nb, 256, 0, s->stream>>>(n, s->uz_ctl.p, s->uz_scal.p, s->uz_q2.p, s->cx.p);
		CK(cudaGetLastError());
		s->launches += 4;
	}
	uz_finish_kernel<<<1, 1, 0, s->stream>>>(s->uz_ctl.p, s->gs_iters_done.p);
	CK(cudaGetLastError());
	s->launches++;
}

void launch_global(S *s)
{
	if (s->linsolver == ADMM_B200_MCGS) launch_mcgs(s);
	else if (s->linsolver == ADMM_B200_UZAWA && !s->obstacles.empty()) launch_uzawa(s);
	else launch_ldlt(s); // LDLT, and UzawaCG with an empty constraint matrix (src/UzawaCG.hpp:78-81)
	launch_mgpu_barrier(s);
}

// ------------------------------------------------------------------------------------------
// finalize helpers
// ------------------------------------------------------------------------------------------
void build_incidence(S *s)
{
	int n = s->n_nodes;
	std::vector<int> off(n + 1, 0);
	size_t slots = 0;
	for (auto t : s->tets) { t->slot_base = slots; slots += (size_t)4 * t->n; for (int i = 0; i < 4 * t->n; ++i) off[t->idx[i] + 1]++; }
	for (auto t : s->tris) { t->slot_base = slots; slots += (size_t)3 * t->n; for (int i = 0; i < 3 * t->n; ++i) off[t->idx[i] + 1]++; }
	for (int i = 0; i < s->pins.n; ++i) off[s->pins.idx[i] + 1]++;
	require(slots < (size_t)0x7fffffff, "too many element corners for int32 slots");
	s->n_slots = slots;
	for (int i = 0; i < n; ++i) off[i + 1] += off[i];
	std::vector<int> fill(off.begin(), off.end() - 1), slot(off[n]);
	for (auto t : s->tets)
		for (int e = 0; e < t->n; ++e)
			for (int c = 0; c < 4; ++c) slot[fill[t->idx[4 * e + c]]++] = (int)(t->slot_base + 4 * (size_t)e + c);
	for (auto t : s->tris)
		for (int e = 0; e < t->n; ++e)
			for (int c = 0; c < 3; ++c) slot[fill[t->idx[3 * e + c]]++] = (int)(t->slot_base + 3 * (size_t)e + c);
	for (int i = 0; i < s->pins.n; ++i) slot[fill[s->pins.idx[i]]++] = (int)(0x80000000u | (unsigned)i);
	// sliced + transposed for the kernel (kernels.cuh, assemble_kernel): warp w owns vertices [32 w, 32 w + 32)
	const int n_warps = (n + 31) / 32;
	std::vector<int> woff(n_warps + 1, 0);
	for (int w = 0; w < n_warps; ++w) {
		int width = 0;
		for (int l = 0; l < 32 && 32 * w + l < n; ++l) width = std::max(width, off[32 * w + l + 1] - off[32 * w + l]);
		woff[w + 1] = woff[w] + width;
	}
	std::vector<int> tslot((size_t)woff[n_warps] * 32, ADMMB200_NO_SLOT);
	for (int i = 0; i < n; ++i)
		for (int k = off[i]; k < off[i + 1]; ++k) tslot[((size_t)woff[i >> 5] + (k - off[i])) * 32 + (i & 31)] = slot[k];
	if (tslot.empty()) tslot.push_back(ADMMB200_NO_SLOT);
	s->inc_off.upload(woff, s->stream);
	s->inc_slot.upload(tslot, s->stream);
	CK(cudaStreamSynchronize(s->stream));
}

template <typename E> void upload_elements(S *s)
{
	const double dt2 = s->dt * s->dt;
	for (auto t : s->tets) {
		t->n_pad = pad32(t->n);
		std::vector<int4> id(t->n);
		for (int e = 0; e < t->n; ++e) id[e] = make_int4(t->idx[4 * e], t->idx[4 * e + 1], t->idx[4 * e + 2], t->idx[4 * e + 3]);
		t->d_idx.upload(id, s->stream);
		CK(cudaStreamSynchronize(s->stream));
		upload_soa<E>(t->d_dminv, t->dminv, t->n, t->n_pad, 9, s->stream);
		std::vector<double> w2(t->n);
		for (int e = 0; e < t->n; ++e) w2[e] = dt2 * t->w[e] * t->w[e];
		upload_soa<E>(t->d_wdt2, w2, t->n, t->n_pad, 1, s->stream);
		t->d_u.alloc((size_t)9 * t->n_pad * sizeof(E)); t->d_u.zero(s->stream);
		// SVD warm start (prox.cuh, svd3_signed): measured on the 1M-tet beam it LOSES 2 us of 60 (the sweeps saved
		// cost less than the extra 32 B/tet and the quaternion conversions), so it is a compile-time option (kernels.cuh)
		if (ADMMB200_SVD_WARMSTART) { t->d_q.alloc((size_t)4 * t->n_pad * sizeof(E)); t->d_q.zero(s->stream); }
		if (s->store_z) { t->d_z.alloc((size_t)9 * t->n_pad * sizeof(E)); t->d_z.zero(s->stream); }
		t->d_defer.alloc((size_t)t->n + 2); t->d_defer.zero(s->stream);
	}
	for (auto t : s->tris) {
		t->n_pad = pad32(t->n);
		std::vector<int4> id(t->n);
		for (int e = 0; e < t->n; ++e) id[e] = make_int4(t->idx[3 * e], t->idx[3 * e + 1], t->idx[3 * e + 2], 0);
		t->d_idx.upload(id, s->stream);
		CK(cudaStreamSynchronize(s->stream));
		upload_soa<E>(t->d_rest, t->rest, t->n, t->n_pad, 4, s->stream);
		std::vector<double> w2(t->n);
		for (int e = 0; e < t->n; ++e) w2[e] = dt2 * t->w[e] * t->w[e];
		upload_soa<E>(t->d_wdt2, w2, t->n, t->n_pad, 1, s->stream);
		t->d_u.alloc((size_t)6 * t->n_pad * sizeof(E)); t->d_u.zero(s->stream);
		if (s->store_z) { t->d_z.alloc((size_t)6 * t->n_pad * sizeof(E)); t->d_z.zero(s->stream); }
	}
	s->f.alloc(std::max<size_t>(s->n_slots, 1) * sizeof(typename Vec4<E>::type));
	s->f.zero(s->stream);
}

void upload_pins(S *s)
{
	PinsH &p = s->pins;
	if (!p.n) return;
	const double dt2 = s->dt * s->dt;
	p.d_idx.upload(p.idx, s->stream);
	p.d_pos.upload(p.pos, s->stream);
	p.d_active.upload(p.active, s->stream);
	std::vector<double> w2(p.n);
	for (int i = 0; i < p.n; ++i) w2[i] = dt2 * p.w[i] * p.w[i];
	p.d_wdt2.upload(w2, s->stream);
	p.d_u.alloc((size_t)3 * p.n); p.d_u.zero(s->stream);
	p.d_z.alloc((size_t)3 * p.n); p.d_z.zero(s->stream);
	p.d_f.alloc(p.n); p.d_f.zero(s->stream);
	CK(cudaStreamSynchronize(s->stream));
}

void upload_gs_pins(S *s)
{
	std::vector<int> slot(std::max(s->n_nodes, 1), -1);
	for (size_t i = 0; i < s->gs_pin_idx.size(); ++i) {
		require(s->gs_pin_idx[i] >= 0 && s->gs_pin_idx[i] < s->n_nodes, "gs pin index out of range");
		slot[s->gs_pin_idx[i]] = (int)i;
	}
	s->d_pin_slot.upload(slot, s->stream);
	std::vector<double> pos = s->gs_pin_pos;
	if (pos.empty()) pos.resize(3, 0.0);
	s->d_pin_pos.upload(pos, s->stream);
	CK(cudaStreamSynchronize(s->stream));
}

int plan_default_lanes()
{
	const char *envl = getenv("ADMM_B200_GS_RES_LANES");
	int lanes = envl ? atoi(envl) : 1;
	if (lanes != 1 && lanes != 2 && lanes != 4) lanes = 1;
	return lanes;
}

// Plans the shared-memory-resident variant and uses it when every part fits one SM's shared memory
// with the chosen value precision; otherwise the streaming kernel (any size) stays in charge.
// ADMM_B200_GS_KERNEL=stream|resident overrides the choice (resident fails loudly if it does not fit).
void build_mcgs_resident(S *s)
{
	const char *env = getenv("ADMM_B200_GS_KERNEL");
	const std::string want = env ? env : "auto";
	s->gs_resident = false;
	if (s->world > 1) require(s->precision == ADMM_B200_FP32 && want != "stream", "multi-GPU needs the resident fp32 Gauss-Seidel (precision FP32)");
	if (want == "stream") { s->gs_info = "stream (forced)"; return; }
	if (s->gs_iters < 2) {
		// The resident kernels stage x_ref (halo included) at the start and write x = x_ref + d at the end of the SAME launch;
		// what keeps a fast part from overwriting positions a slow neighbour has not staged yet is that it has to wait for that
		// neighbour's published values first -- guaranteed only when every pair of neighbours exchanges at least once in each
		// direction, i.e. from the second sweep on.
		require(s->world == 1 && want != "resident", "the resident Gauss-Seidel needs at least 2 sweeps per solve");
		s->gs_info = "stream (fewer than 2 sweeps per solve)";
		return;
	}
	const int val_bytes = s->precision == ADMM_B200_FP64 ? 8 : 4;
	int max_optin = 0;
	CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
	const bool prof = getenv("ADMM_B200_GS_PROF") != nullptr;
	s->gs_tiled = false;
	// The tiled fp32 solve (mcgs_tiled_f32.cuh: short tasks of 8 nodes x 4 lanes) is an ALTERNATIVE, selected with
	// ADMM_B200_GS_TILED=1: measured on the 1M-tet beam it needs 0.59 ms per solve where the static-ownership kernel needs
	// 0.36 (profiles/r02c_gsprof2.log: the shared-memory pipe, not the task latency, bounds a pass, and four lanes per node
	// cost more bank conflicts per gathered float4 than one).  Needs equal x/y/z masses per node and no obstacles.
	bool want_tiled = false;
	if (const char *et = getenv("ADMM_B200_GS_TILED")) want_tiled = atoi(et) != 0 && val_bytes == 4 && s->obstacles.empty();
	if (getenv("ADMM_B200_GS_OWNED") || getenv("ADMM_B200_GS_RES_LANES")) want_tiled = false; // a round-1 variant was asked for by name
	for (int i = 0; i < s->n_nodes && want_tiled; ++i) want_tiled = s->h_m[3 * (size_t)i] == s->h_m[3 * (size_t)i + 1] && s->h_m[3 * (size_t)i] == s->h_m[3 * (size_t)i + 2];
	ResidentPlan R;
	int lanes = 1;
	size_t need = 0, budget = 0;
	const void *kern = nullptr;
	for (int attempt = want_tiled ? 0 : 1; attempt < 2; ++attempt) {
		const bool tiled = attempt == 0;
		lanes = tiled ? 4 : plan_default_lanes();
		kern = tiled ? (prof ? (const void *)mcgs_tiled_f32_kernel<4, true> : (const void *)mcgs_tiled_f32_kernel<4, false>) : resident_kernel_ptr(val_bytes == 8, lanes, prof);
		cudaFuncAttributes fa;
		CK(cudaFuncGetAttributes(&fa, kern));
		budget = (size_t)max_optin - fa.sharedSizeBytes;
		try {
			R = plan_resident(s->n_nodes, s->L_rowptr.data(), s->L_cols.data(), s->L_vals.data(), s->n_colors, s->color_off.data(), s->color_nodes.data(), s->h_x0.data(), s->gs_parts * s->world, lanes);
		} catch (std::exception &e) {
			if (want == "resident" || s->world > 1) throw;
			s->gs_info = std::string("stream (") + e.what() + ")";
			return;
		}
		need = R.smem_bytes(s->n_colors, val_bytes, val_bytes == 8 ? 0 : (tiled ? 2 : 1));
		if (tiled && need <= budget && R.max_own <= 65535) { s->gs_tiled = true; break; }
	}
	s->gs_res_lanes = lanes;
	char buf[256];
	snprintf(buf, sizeof(buf), "%d lane(s)/node, %zu B shared memory per part needed (max own %zu, halo %zu, rows %zu, neighbours %zu; ELL fill %.3f), budget %zu B", lanes, need, R.max_own, R.max_halo, R.max_rows,
		R.max_nbr, R.entries ? (double)R.nnz / (double)R.entries : 1.0, budget);
	if (need > budget) {
		if (want == "resident" || s->world > 1) throw std::runtime_error(std::string("resident MCGS does not fit: ") + buf);
		s->gs_info = std::string("stream: ") + buf;
		return;
	}
	s->res_parts.upload(R.parts, s->stream);
	s->res_col.upload(R.col, s->stream);
	s->res_gid.upload(R.gid, s->stream);
	s->res_slice_row.upload(R.slice_row, s->stream);
	s->res_color_slice.upload(R.color_slice, s->stream);
	s->res_slice_node.upload(R.slice_node, s->stream);
	s->res_nbr.upload(R.nbr.empty() ? std::vector<int>(1, 0) : R.nbr, s->stream);
	s->res_halo_color.upload(R.halo_color, s->stream);
	if (s->world > 1) {
		s->mg_dest_mask.upload(dest_masks(R, s->n_nodes, s->gs_parts, s->rank), s->stream);
		s->mg_flags.alloc(8 * ADMMB200_MAX_RANKS); s->mg_flags.zero(s->stream);
		// nodes this rank owns, and owned + ghost nodes (halo nodes of its parts that another rank owns)
		std::vector<char> loc((size_t)s->n_nodes, 0);
		for (int i = 0; i < s->n_nodes; ++i) if (R.part_of[i] / s->gs_parts == s->rank) loc[i] = 1;
		for (int p = s->rank * s->gs_parts; p < (s->rank + 1) * s->gs_parts; ++p) {
			const PartDesc &d = R.parts[p];
			for (int h = 0; h < d.n_halo; ++h) { const int g = R.gid[d.gid_off + d.n_own + h]; if (!loc[g]) loc[g] = 2; }
		}
		s->mg_local.clear(); s->mg_owned.clear();
		for (int i = 0; i < s->n_nodes; ++i) { if (loc[i]) s->mg_local.push_back(i); if (loc[i] == 1) s->mg_owned.push_back(i); }
		s->mg_runs.clear();
		for (size_t i = 0; i < s->mg_local.size();) {
			size_t j = i + 1;
			while (j < s->mg_local.size() && s->mg_local[j] == s->mg_local[j - 1] + 1) ++j;
			s->mg_runs.push_back(s->mg_local[i]); s->mg_runs.push_back((int)(j - i)); s->mg_runs.push_back((int)i);
			i = j;
		}
		s->d_mg_local.upload(s->mg_local.empty() ? std::vector<int>(1, 0) : s->mg_local, s->stream);
		s->d_mg_owned.upload(s->mg_owned.empty() ? std::vector<int>(1, 0) : s->mg_owned, s->stream);
		if (s->mg_pinned) { cudaFreeHost(s->mg_pinned); s->mg_pinned = nullptr; }
		CK(cudaMallocHost((void **)&s->mg_pinned, sizeof(double) * 6 * std::max<size_t>(s->mg_local.size(), 1)));
	}
	if (val_bytes == 4) {
		require((long long)s->gs_iters * s->n_colors < 4094, "resident fp32 MCGS: sweeps x colours must stay below 4094 (12-bit pass tags)");
		plan_mailboxes(R, s->n_nodes, s->gs_parts);
		s->res_parts.upload(R.parts, s->stream); // again: now with the mailbox offsets
		s->res_dest_off.upload(R.dest_off, s->stream);
		s->res_dest_slot.upload(R.dest_slot, s->stream);
		s->res_total_slots = (int)R.total_slots;
		// published increments: [2][n_nodes][3] words by node id (table-walking kernel) or [2][3][slots] (static-ownership kernel)
		s->res_dglob.alloc(6 * std::max((size_t)s->n_nodes, R.total_slots)); s->res_dglob.zero(s->stream);
		s->res_nodebuf.alloc(2 * (size_t)s->n_nodes); s->res_nodebuf.zero(s->stream);
	}
	s->res_sync.alloc(8 * (size_t)s->gs_parts + 2 * (size_t)std::max(s->gs_iters, 1));
	s->res_scratch_resid_n = 2 * (size_t)s->gs_iters + 4;
	s->res_scratch.alloc(s->res_scratch_resid_n + (2 + 8 * (size_t)s->gs_parts + 2 * (size_t)std::max(s->gs_iters, 1) + 1) / 2 + 1);
	require(R.max_nbr <= 192, "resident plan: too many neighbour parts");
	if (prof) { s->res_prof.alloc(16 * (size_t)s->gs_parts + 1024 + 128 * (size_t)s->gs_parts); s->res_prof.zero(s->stream); }
	if (val_bytes == 8) {
		s->res_val.alloc(std::max<size_t>(R.val.size(), 1) * 8);
		if (!R.val.empty()) CK(cudaMemcpyAsync(s->res_val.p, R.val.data(), R.val.size() * 8, cudaMemcpyHostToDevice, s->stream));
		CK(cudaStreamSynchronize(s->stream));
	} else {
		s->res_val64.upload(R.val.empty() ? std::vector<double>(1, 0.0) : R.val, s->stream);
		std::vector<float> v32(R.val.begin(), R.val.end());
		s->res_val.alloc(std::max<size_t>(v32.size(), 1) * 4);
		if (!v32.empty()) CK(cudaMemcpyAsync(s->res_val.p, v32.data(), v32.size() * 4, cudaMemcpyHostToDevice, s->stream));
		CK(cudaStreamSynchronize(s->stream));
	}
	CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
	if (s->gs_tiled) {
		// a'_ij = omega a_ij / a_ii: the row's node is the one the entry's lane belongs to (T lanes per node, G nodes per slice)
		const int T = lanes, G = 32 / T;
		std::vector<float> vs(std::max<size_t>(R.val.size(), 1), 0.f);
		std::vector<uint4> d4((size_t)s->n_nodes, make_uint4(0u, 0u, 0u, 0u));
		for (const PartDesc &d : R.parts) {
			const int *gid = R.gid.data() + d.gid_off;
			for (int sl = 0; sl < d.n_slices; ++sl)
				for (int g = 0; g < G; ++g) {
					const int l = R.slice_node[d.snode_off + sl * G + g];
					if (l < 0) continue;
					const int node = gid[l];
					double aii = s->h_m[3 * (size_t)node];
					for (int q = s->L_rowptr[node]; q < s->L_rowptr[node + 1]; ++q) if (s->L_cols[q] == node) aii += s->L_vals[q];
					const double sc = s->gs_omega / aii;
					for (int r = R.slice_row[d.slice_off + sl]; r < R.slice_row[d.slice_off + sl + 1]; ++r)
						for (int t = 0; t < T; ++t) { const size_t e = (size_t)d.ent_off + (size_t)r * 32 + g * T + t; vs[e] = (float)(sc * R.val[e]); }
				}
			for (int l = 0; l < d.n_own; ++l) {
				const int e0 = R.dest_off[d.own_off + l], cnt = std::min(R.dest_off[d.own_off + l + 1] - e0, 63);
				d4[(size_t)d.own_off + l] = make_uint4(cnt > 0 ? R.dest_slot[e0] : 0u, cnt > 1 ? R.dest_slot[e0 + 1] : 0u, (unsigned int)e0, (unsigned int)cnt);
			}
		}
		s->res_val_scaled.upload(vs, s->stream);
		s->res_dest4.upload(d4, s->stream);
		CK(cudaStreamSynchronize(s->stream));
	}
	// static-ownership kernel (mcgs_owned_f32.cuh): fp32 sweeps, one lane per node, every part's slices fit the warps' registers
	s->gs_owned_threads = 0;
	if (!s->gs_tiled && val_bytes == 4 && lanes == 1 && s->n_colors <= ADMMB200_OWNED_MAX_COLORS && (long long)s->gs_iters * s->n_colors <= 510) { // 9-bit pass tags in its mailbox words
		const char *eo = getenv("ADMM_B200_GS_OWNED"); // 0: keep the table-walking kernel; 512 / 768: force that variant
		const int want_threads = eo ? atoi(eo) : -1;
		int pick = 0;
		if (want_threads == 512 || want_threads == 768) pick = want_threads;
		else if (want_threads != 0) pick = (int)R.max_slices <= owned_capacity(512) ? 512 : 768;
		if (pick && (int)R.max_slices <= owned_capacity(pick)) {
			for (int ob = 0; ob < 2; ++ob) {
				const void *ko = owned_kernel_ptr(pick, ob != 0, prof);
				cudaFuncAttributes fo;
				CK(cudaFuncGetAttributes(&fo, ko));
				if (need + fo.sharedSizeBytes > (size_t)max_optin) { pick = 0; break; }
				CK(cudaFuncSetAttribute(ko, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
			}
			s->gs_owned_threads = pick;
		}
	}
	CK(cudaStreamSynchronize(s->stream));
	s->gs_res_smem = need;
	s->gs_resident = true;
	s->gs_info = std::string("resident: ") + buf + (s->gs_tiled ? std::string("; tiled kernel (8 nodes x 4 lanes per task), 512 threads") :
		s->gs_owned_threads ? "; static-ownership kernel, " + std::to_string(s->gs_owned_threads) + " threads" : std::string("; table-walking kernel"));
}

void build_mcgs(S *s)
{
	const int n = s->n_nodes;
	require(s->sys_n == n, "set_system: matrix size does not match the node count");
	require(s->n_colors > 0, "NodalMultiColorGS needs colours (admm_b200_set_colors)");
	const char *env = getenv("ADMM_B200_GS_LANES");
	int T = env ? atoi(env) : 4;
	if (T != 1 && T != 2 && T != 4 && T != 8) T = 4;
	s->gs_lanes = T;
	const int G = 32 / T;
	std::vector<double> diag((size_t)3 * n, 0.0);
	std::vector<char> seen(n, 0);
	std::vector<int> color_first(s->n_colors + 1, 0), slice_ptr(1, 0), slice_node, ell_col;
	std::vector<double> ell_val;
	size_t nnz = 0;
	for (int c = 0; c < s->n_colors; ++c) {
		int k0 = s->color_off[c], k1 = s->color_off[c + 1];
		for (int k = k0; k < k1; k += G) {
			int width = 0;
			int nodes[32];
			for (int g = 0; g < G; ++g) {
				int node = (k + g < k1) ? s->color_nodes[k + g] : -1;
				nodes[g] = node;
				if (node < 0) continue;
				require(node < n, "colour list: node out of range");
				require(!seen[node], "colour list: node appears twice");
				seen[node] = 1;
				int len = 0;
				for (int q = s->L_rowptr[node]; q < s->L_rowptr[node + 1]; ++q)
					if (s->L_cols[q] != node && s->L_vals[q] != 0.0) len++;
				width = std::max(width, (len + T - 1) / T);
			}
			size_t base = ell_col.size();
			ell_col.resize(base + (size_t)width * 32, 0);
			ell_val.resize(base + (size_t)width * 32, 0.0);
			for (int g = 0; g < G; ++g) {
				int node = nodes[g];
				slice_node.push_back(node);
				if (node < 0) continue;
				int j = 0;
				for (int q = s->L_rowptr[node]; q < s->L_rowptr[node + 1]; ++q) {
					int col = s->L_cols[q];
					double val = s->L_vals[q];
					if (col == node) { for (int d = 0; d < 3; ++d) diag[3 * (size_t)node + d] += val; continue; }
					if (val == 0.0) continue;
					require(col >= 0 && col < n, "set_system: column out of range");
					int r = j / T, t = j % T;
					ell_col[base + (size_t)r * 32 + g * T + t] = col;
					ell_val[base + (size_t)r * 32 + g * T + t] = val;
					++j; ++nnz;
				}
				// padding entries read the node itself with a zero coefficient
				for (; j < width * T; ++j) { int r = j / T, t = j % T; ell_col[base + (size_t)r * 32 + g * T + t] = node; }
			}
			for (int g = 0; g < G; ++g) if (nodes[g] < 0) for (int r = 0; r < width; ++r) for (int t = 0; t < T; ++t) ell_col[base + (size_t)r * 32 + g * T + t] = nodes[0];
			slice_ptr.push_back((int)(ell_col.size() / 32));
		}
		color_first[c + 1] = (int)slice_ptr.size() - 1;
	}
	for (int i = 0; i < n; ++i) require(seen[i], "colour list: a node has no colour");
	for (int i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) {
		diag[3 * (size_t)i + d] += s->h_m[3 * (size_t)i + d];
		// LinearSolver::is_zero (src/LinearSolver.hpp:51), "Zero on diagonal" (src/NodalMultiColorGS.hpp:203-206)
		require(std::abs(diag[3 * (size_t)i + d]) >= 2.2250738585072014e-308, "**NodalMultiColorGS Error: Zero on diagonal");
	}
	s->gs_nnz = nnz; s->gs_ell_entries = ell_col.size();
	s->gs_color_first.upload(color_first, s->stream);
	s->gs_slice_ptr.upload(slice_ptr, s->stream);
	s->gs_slice_node.upload(slice_node, s->stream);
	s->gs_ell_col.upload(ell_col, s->stream);
	s->gs_ell_val.upload(ell_val, s->stream);
	s->gs_diag.upload(diag, s->stream);
	s->gs_resid.alloc(2 * (size_t)s->gs_iters + 4); // [0] |b|^2, [1..iters] residuals, then iters lower bounds
	s->gs_iters_done.alloc(1);
	CK(cudaStreamSynchronize(s->stream));
	int occ = 1;
	switch (T) { case 1: occ = mcgs_occupancy<1>(); break; case 2: occ = mcgs_occupancy<2>(); break; case 8: occ = mcgs_occupancy<8>(); break; default: occ = mcgs_occupancy<4>(); }
	require(occ >= 1, "mcgs kernel does not fit on an SM");
	const char *envb = getenv("ADMM_B200_GS_BLOCKS_PER_SM");
	int bps = envb ? std::max(1, std::min(occ, atoi(envb))) : 1;
	s->gs_grid = s->n_sms * bps;
	if (!s->obstacles.empty()) { s->d_obstacles.upload(s->obstacles, s->stream); CK(cudaStreamSynchronize(s->stream)); }
	upload_gs_pins(s);
	build_mcgs_resident(s);
}

void build_ldlt(S *s)
{
	require(s->have_ldlt, "LDLT / UzawaCG need a factor (admm_b200_set_ldlt)");
	const int n = s->ld_n;
	require(n == s->n_nodes, "set_ldlt: factor size does not match the node count");
	for (int i = 0; i < n; ++i) require(s->h_m[3 * (size_t)i] == s->h_m[3 * (size_t)i + 1] && s->h_m[3 * (size_t)i] == s->h_m[3 * (size_t)i + 2], "LDLT path needs equal x/y/z masses per node");
	const std::vector<int> &Lp = s->ld_Lp, &Li = s->ld_Li;
	const std::vector<double> &Lx = s->ld_Lx;
	for (int j = 0; j < n; ++j) for (int q = Lp[j]; q < Lp[j + 1]; ++q) require(Li[q] > j && Li[q] < n && (q == Lp[j] || Li[q] > Li[q - 1]), "set_ldlt: L must be strictly lower, CSC, ascending rows");
	{
		const char *ek = getenv("ADMM_B200_LDLT_KERNEL");
		s->ld_blocks = !(ek && std::string(ek) == "levels");
	}
	if (s->ld_blocks) {
		LdltBlockPlan B = plan_ldlt_blocks(n, Lp.data(), Li.data(), Lx.data(), s->n_sms, 1024);
		s->lb_levels = B.n_levels_f; s->lb_cut = B.cut;
		B.build_descriptors(s->ld_perm.data());
		auto up4 = [&](DevBuf<int4> &dst, const std::vector<int> &src) { dst.alloc(src.size() / 4); CK(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream)); CK(cudaStreamSynchronize(s->stream)); };
		up4(s->lb_desc_fg, B.desc_fg); up4(s->lb_desc_bg, B.desc_bg); up4(s->lb_desc_d, B.desc_d);
		s->lb_desc_off.upload(B.desc_off, s->stream);
		s->d_ld_perm.upload(s->ld_perm, s->stream);
		s->lb_blk_of.upload(B.blk_of, s->stream); s->lb_blk_c0.upload(B.blk_c0, s->stream); s->lb_inv_off.upload(B.inv_off, s->stream);
		s->lb_inv.upload(B.inv, s->stream); s->lb_invT.upload(B.invT, s->stream);
		auto nonempty_i = [](std::vector<int> &v) { if (v.empty()) v.push_back(0); };
		auto nonempty_d = [](std::vector<double> &v) { if (v.empty()) v.push_back(0.0); };
		nonempty_i(B.f_cols); nonempty_d(B.f_vals); nonempty_i(B.b_rows); nonempty_d(B.b_vals);
		s->lb_lev_ptr.upload(B.f_lev_ptr, s->stream); s->lb_rows.upload(B.f_rows, s->stream); s->lb_lanes.upload(B.lanes, s->stream);
		s->lb_f_rowptr.upload(B.f_rowptr, s->stream); s->lb_f_cols.upload(B.f_cols, s->stream); s->lb_f_vals.upload(B.f_vals, s->stream);
		s->lb_b_colptr.upload(B.b_colptr, s->stream); s->lb_b_rows.upload(B.b_rows, s->stream); s->lb_b_vals.upload(B.b_vals, s->stream);
		s->lb_seg_ptr.upload(B.seg_ptr, s->stream); s->lb_seg_begin.upload(B.seg_begin, s->stream); s->lb_seg_end.upload(B.seg_end, s->stream); s->lb_seg_level.upload(B.seg_level, s->stream);
		s->d_ld_D.upload(s->ld_D, s->stream);
		s->d_ld_y.alloc(n); s->lb_t.alloc(n);
		if (getenv("ADMM_B200_LDLT_PROF")) { s->lb_prof.alloc(4 * (size_t)B.n_levels_f + 8); s->lb_prof.zero(s->stream); }
		CK(cudaStreamSynchronize(s->stream));
		int occ = 0;
		CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ldlt_blocks_kernel, 1024, 0));
		require(occ >= 1, "ldlt block kernel does not fit on an SM");
		s->ld_grid = s->n_sms;
		char buf[320];
		snprintf(buf, sizeof(buf), "ldlt blocks: n %d, nnz(L) %lld, %d blocks (largest %d), %lld entries outside + %lld in the inverted diagonal blocks, %d levels of which %d in the CTA-local bottom forest (%d segments), %d CTAs x 1024", n, (long long)Lp[n],
			B.n_blocks, B.max_block, B.nnz_out, B.nnz_inv, B.n_levels_f, B.cut, (int)B.seg_ptr.back(), s->ld_grid);
		s->gs_info = buf;
	} else {
	// CSR of strictly lower L (row gather for the forward solve)
	std::vector<int> rp(n + 1, 0);
	for (int j = 0; j < n; ++j) for (int q = Lp[j]; q < Lp[j + 1]; ++q) { require(Li[q] > j && Li[q] < n, "set_ldlt: L must be strictly lower, CSC"); rp[Li[q] + 1]++; }
	for (int i = 0; i < n; ++i) rp[i + 1] += rp[i];
	std::vector<int> fill(rp.begin(), rp.end() - 1), rc(Lp[n]);
	std::vector<double> rv(Lp[n]);
	for (int j = 0; j < n; ++j) for (int q = Lp[j]; q < Lp[j + 1]; ++q) { int i = Li[q]; rc[fill[i]] = j; rv[fill[i]] = Lx[q]; fill[i]++; }
	// levels
	std::vector<int> lf(n, 0), lb(n, 0);
	int nlf = 0, nlb = 0;
	for (int i = 0; i < n; ++i) { int l = 0; for (int q = rp[i]; q < rp[i + 1]; ++q) l = std::max(l, lf[rc[q]] + 1); lf[i] = l; nlf = std::max(nlf, l + 1); }
	for (int j = n - 1; j >= 0; --j) { int l = 0; for (int q = Lp[j]; q < Lp[j + 1]; ++q) l = std::max(l, lb[Li[q]] + 1); lb[j] = l; nlb = std::max(nlb, l + 1); }
	auto bucket = [&](const std::vector<int> &lev, int nl, std::vector<int> &ptr, std::vector<int> &rows) {
		ptr.assign(nl + 1, 0);
		for (int i = 0; i < n; ++i) ptr[lev[i] + 1]++;
		for (int l = 0; l < nl; ++l) ptr[l + 1] += ptr[l];
		std::vector<int> f2(ptr.begin(), ptr.end() - 1);
		rows.resize(n);
		for (int i = 0; i < n; ++i) rows[f2[lev[i]]++] = i;
	};
	std::vector<int> fptr, frows, bptr, brows;
	bucket(lf, nlf, fptr, frows);
	bucket(lb, nlb, bptr, brows);
	s->ld_levels_fwd = nlf; s->ld_levels_bwd = nlb;
	s->d_ld_perm.upload(s->ld_perm, s->stream);
	s->d_fwd_level_ptr.upload(fptr, s->stream); s->d_fwd_rows.upload(frows, s->stream);
	s->d_fwd_rowptr.upload(rp, s->stream); s->d_fwd_cols.upload(rc, s->stream); s->d_fwd_vals.upload(rv, s->stream);
	s->d_bwd_level_ptr.upload(bptr, s->stream); s->d_bwd_rows.upload(brows, s->stream);
	s->d_bwd_rowptr.upload(s->ld_Lp, s->stream); s->d_bwd_cols.upload(s->ld_Li, s->stream); s->d_bwd_vals.upload(s->ld_Lx, s->stream);
	s->d_ld_D.upload(s->ld_D, s->stream);
	s->d_ld_y.alloc(n);
	} // level-scheduled structures
	if (s->linsolver == ADMM_B200_UZAWA && !s->obstacles.empty()) {
		// UzawaCG with passive collisions (uzawa.cuh): at most one constraint row per node
		s->uz_hv.alloc(n); s->uz_hn.alloc(3 * (size_t)n); s->uz_hc.alloc(n); s->uz_y.alloc(n); s->uz_r.alloc(n); s->uz_d.alloc(n); s->uz_q3.alloc(n);
		s->uz_q1.alloc(n); s->uz_q2.alloc(n); s->uz_ctl.alloc(8); s->uz_scal.alloc(2);
		s->uz_y.zero(s->stream); s->uz_ctl.zero(s->stream); s->uz_scal.zero(s->stream);
		s->uz_q1.zero(s->stream); // uzawa_blocks.cuh keeps it all zero between its scatters
		s->uz_cta_cnt.alloc(std::max(s->n_sms, 1)); s->uz_red.alloc(4 * (size_t)std::max(s->n_sms, 1));
		for (int v : s->surface_inds) require(v >= 0 && v < n, "set_surface_inds: vertex out of range");
		if (!s->surface_inds.empty()) s->uz_cand.upload(s->surface_inds, s->stream);
		if (!s->gs_iters_done.p) s->gs_iters_done.alloc(1);
		s->d_obstacles.upload(s->obstacles, s->stream);
		CK(cudaStreamSynchronize(s->stream));
	}
	CK(cudaStreamSynchronize(s->stream));
	if (s->ld_blocks) return;
	const char *env = getenv("ADMM_B200_LDLT_LANES");
	int T = env ? atoi(env) : 4;
	if (T != 1 && T != 2 && T != 4 && T != 8) T = 4;
	s->ld_lanes = T;
	s->ld_grid = s->n_sms;
	char buf[256];
	snprintf(buf, sizeof(buf), "ldlt: n %d, nnz(L) %lld, dependency levels %d forward + %d backward, %d lane(s)/row, %d CTAs", n, (long long)Lp[n], s->ld_levels_fwd, s->ld_levels_bwd, T, s->ld_grid);
	s->gs_info = buf;
}

// Events tightly around the hot kernels (timers on): kind 0 tet prox, 1 assemble, 2 solve.  Summed after the step.
void fine_begin(S *s, int kind)
{
	if (!s->fine_on) return;
	cudaEvent_t a, b;
	s->fine_used = 2 * s->fine_kind.size();
	if (s->fine_pool.size() < s->fine_used + 2) { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); s->fine_pool.push_back(a); s->fine_pool.push_back(b); }
	a = s->fine_pool[s->fine_used];
	CK(cudaEventRecord(a, s->stream));
	s->fine_kind.push_back(kind);
}
void fine_end(S *s)
{
	if (!s->fine_on) return;
	CK(cudaEventRecord(s->fine_pool[s->fine_used + 1], s->stream));
	s->fine_used += 2;
}

cudaEvent_t get_event(S *s, size_t i)
{
	while (s->events.size() <= i) { cudaEvent_t e; CK(cudaEventCreate(&e)); s->events.push_back(e); }
	return s->events[i];
}

void zero_duals(S *s)
{
	for (auto t : s->tets) t->d_u.zero(s->stream);
	for (auto t : s->tris) t->d_u.zero(s->stream);
	s->pins.d_u.zero(s->stream);
}

// Timers.  Immediate mode (a runtime pointer is passed): events of this step only, summed after a stream
// synchronise at the end of the step -- what Solver::runtime_data() needs.  Deferred mode
// (admm_b200_set_deferred_timers): every step records into fresh events and returns WITHOUT synchronising, so
// the host can queue the next step while this one runs; admm_b200_collect_timers sums all steps since the
// last collection.  One block of events per step: 4 per ADMM iteration + 1, plus 2 per hot-kernel launch.
void sum_timer_blocks(S *s, admm_b200_runtime *rt)
{
	CK(cudaStreamSynchronize(s->stream));
	rt->global_ms = rt->local_ms = rt->collision_ms = rt->assemble_ms = rt->step_ms = 0; rt->inner_iters = 0;
	for (int k = 0; k < 3; ++k) { s->kernel_ms[k] = 0; s->kernel_n[k] = 0; }
	const bool logged = s->linsolver == ADMM_B200_MCGS || (s->linsolver == ADMM_B200_UZAWA && !s->obstacles.empty());
	for (const TimerBlock &tb : s->tblocks) {
		for (int it = 0; it < tb.iters; ++it) {
			float a = 0, b = 0, c = 0;
			CK(cudaEventElapsedTime(&a, s->events[tb.ev0 + 4 * it], s->events[tb.ev0 + 4 * it + 1]));
			CK(cudaEventElapsedTime(&b, s->events[tb.ev0 + 4 * it + 1], s->events[tb.ev0 + 4 * it + 2]));
			CK(cudaEventElapsedTime(&c, s->events[tb.ev0 + 4 * it + 2], s->events[tb.ev0 + 4 * it + 3]));
			rt->local_ms += a; rt->assemble_ms += b; rt->global_ms += b + c;
		}
		if (tb.iters > 0) { float t = 0; CK(cudaEventElapsedTime(&t, s->events[tb.ev0], s->events[tb.ev0 + 4 * tb.iters])); rt->step_ms += t; }
		for (size_t i = tb.fine0; i < tb.fine1; ++i) {
			float t = 0;
			CK(cudaEventElapsedTime(&t, s->fine_pool[2 * i], s->fine_pool[2 * i + 1]));
			s->kernel_ms[s->fine_kind[i]] += t; s->kernel_n[s->fine_kind[i]]++;
		}
		if (logged) {
			std::vector<int> its(tb.iters);
			if (tb.iters) CK(cudaMemcpy(its.data(), s->iter_log.p + tb.log0, sizeof(int) * tb.iters, cudaMemcpyDeviceToHost));
			for (int i : its) rt->inner_iters += i;
		} else rt->inner_iters += tb.iters; // LDLT / empty-C Uzawa return 1 per solve
	}
	s->timer_steps = (int)s->tblocks.size();
	s->tblocks.clear(); s->ev_next = 0; s->fine_used = 0; s->fine_kind.clear(); s->log_next = 0;
}

void do_step(S *s, int admm_iters, double gravity, admm_b200_runtime *rt)
{
	require(s->finalized, "step before finalize");
	require(admm_iters >= 0 && admm_iters <= (int)s->iter_log.n, "admm_iters out of range");
	// deferred mode may sample: events on every deferred_stride-th step only (an event costs ~3 us of stream time and a
	// step records 10 per ADMM iteration: 0.64 ms of an 8.9 ms step, tools/timer_overhead.py)
	const bool timed = rt != nullptr || (s->deferred_timers && (s->deferred_count++ % s->deferred_stride) == 0);
	if (timed) {
		// immediate mode starts from a clean slate; deferred mode appends (and folds into the pending sums when the
		// iteration log or the event pools would grow without bound)
		if (rt && !s->tblocks.empty()) { admm_b200_runtime tmp; sum_timer_blocks(s, &tmp); }
		if (!rt && (s->log_next + admm_iters > (int)s->iter_log.n || s->tblocks.size() >= 256)) {
			admm_b200_runtime tmp; double km[3]; long long kn[3];
			sum_timer_blocks(s, &tmp);
			for (int k = 0; k < 3; ++k) { km[k] = s->kernel_ms[k]; kn[k] = s->kernel_n[k]; }
			s->pend.global_ms += tmp.global_ms; s->pend.local_ms += tmp.local_ms; s->pend.assemble_ms += tmp.assemble_ms; s->pend.step_ms += tmp.step_ms; s->pend.inner_iters += tmp.inner_iters;
			for (int k = 0; k < 3; ++k) { s->pend_kernel_ms[k] += km[k]; s->pend_kernel_n[k] += kn[k]; }
			s->pend_steps += s->timer_steps;
		}
	}
	const int n = s->n_nodes;
	const int nb = (n + 255) / 256;
	for (auto &wp : s->winds) {
		S::Wind &w = *wp;
		const int nt = (int)(w.tris.size() / 3);
		if (nt == 0) continue;
		if (!w.uploaded) {
			// node -> its triangles, in triangle order (the order the kicks are added in)
			std::vector<int> cnt((size_t)n, 0), nodes, ptr(1, 0), inc;
			for (int t = 0; t < 3 * nt; ++t) { require(w.tris[t] >= 0 && w.tris[t] < n, "wind force: vertex index out of range"); cnt[w.tris[t]]++; }
			std::vector<int> slot((size_t)n, -1);
			for (int i = 0; i < n; ++i) if (cnt[i]) { slot[i] = (int)nodes.size(); nodes.push_back(i); ptr.push_back(ptr.back() + cnt[i]); }
			inc.resize((size_t)ptr.back());
			std::vector<int> fill(ptr.begin(), ptr.end() - 1);
			for (int t = 0; t < nt; ++t) for (int c = 0; c < 3; ++c) inc[(size_t)fill[slot[w.tris[3 * t + c]]]++] = t;
			w.n_touched = (int)nodes.size();
			w.d_tris.upload(w.tris, s->stream); w.d_nodes.upload(nodes, s->stream); w.d_inc_ptr.upload(ptr, s->stream); w.d_inc_tri.upload(inc, s->stream);
			w.d_kick.alloc((size_t)nt);
			CK(cudaStreamSynchronize(s->stream)); // the host vectors above go out of scope
			w.uploaded = true;
		}
		wind_tri_kernel<<<(nt + 255) / 256, 256, 0, s->stream>>>(nt, w.d_tris.p, w.dir[0], w.dir[1], w.dir[2], s->dt, s->x.p, s->v.p, w.d_kick.p);
		wind_node_kernel<<<(w.n_touched + 255) / 256, 256, 0, s->stream>>>(w.n_touched, w.d_nodes.p, w.d_inc_ptr.p, w.d_inc_tri.p, w.d_kick.p, s->v.p);
		CK(cudaGetLastError());
		s->launches += 2;
	}
	step_begin_kernel<<<nb, 256, 0, s->stream>>>(n, s->dt, gravity, s->x.p, s->v.p, s->m.p, s->mxbar.p, s->cx.p);
	CK(cudaGetLastError());
	s->launches++;
	zero_duals(s); // curr_u = 0 every step (src/Solver.cpp:71)
	TimerBlock tb;
	tb.ev0 = s->ev_next; tb.iters = admm_iters; tb.fine0 = s->fine_kind.size(); tb.log0 = s->log_next;
	size_t ev = s->ev_next;
	s->fine_on = timed;
	const bool logged = s->linsolver == ADMM_B200_MCGS || (s->linsolver == ADMM_B200_UZAWA && !s->obstacles.empty());
	// events per ADMM iteration: [4it] local [4it+1] assemble [4it+2] solve [4it+3]
	for (int it = 0; it < admm_iters; ++it) {
		if (timed) CK(cudaEventRecord(get_event(s, ev++), s->stream));
		launch_local(s);
		if (timed) CK(cudaEventRecord(get_event(s, ev++), s->stream));
		launch_assemble(s);
		if (timed) CK(cudaEventRecord(get_event(s, ev++), s->stream));
		launch_global(s);
		if (timed) {
			CK(cudaEventRecord(get_event(s, ev++), s->stream));
			// inner_iters += solve() (src/Solver.cpp:99): read back when the timers are summed
			if (logged) CK(cudaMemcpyAsync(s->iter_log.p + tb.log0 + it, s->gs_iters_done.p, sizeof(int), cudaMemcpyDeviceToDevice, s->stream));
		}
	}
	step_end_kernel<<<nb, 256, 0, s->stream>>>(n, s->dt, s->x.p, s->v.p, s->cx.p);
	CK(cudaGetLastError());
	s->launches++;
	s->fine_on = false;
	if (timed) {
		CK(cudaEventRecord(get_event(s, ev++), s->stream)); // end of the step: events[ev0 + 4 iters]
		tb.fine1 = s->fine_kind.size();
		s->ev_next = ev; s->log_next += admm_iters;
		s->tblocks.push_back(tb);
		if (rt) sum_timer_blocks(s, rt);
	}
}

__global__ void scatter3_to4_kernel(int n, const int *__restrict__ idx, const double *__restrict__ in3, double4 *__restrict__ out4)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	st_node(&out4[idx[i]], in3[3 * i], in3[3 * i + 1], in3[3 * i + 2]);
}
__global__ void gather4_to3_kernel(int n, const int *__restrict__ idx, const double4 *__restrict__ in4, double *__restrict__ out3)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double4 v = in4[idx[i]];
	out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
}

// Multi-GPU variants of upload_state / download_state: a rank computes on its owned nodes and reads its ghost nodes, so
// only those travel, in both directions.  Host arrays keep the full 3n layout: entries of nodes this rank neither owns nor
// reads are left untouched (a caller merges the ranks' results by Solver::node_owner()).
void upload_state_local(S *s, const double *x, const double *v)
{
	const int nl = (int)s->mg_local.size();
	if (!nl) return;
	if (s->mg_runs.size() <= 3 * 64) {
		// a spatially compact rank of a mesh numbered along an axis is a handful of id ranges: copy them where they lie
		for (size_t r = 0; r < s->mg_runs.size(); r += 3) {
			const size_t first = (size_t)s->mg_runs[r], len = (size_t)s->mg_runs[r + 1], off = (size_t)s->mg_runs[r + 2];
			CK(cudaMemcpyAsync(s->stage3.p + 3 * off, x + 3 * first, sizeof(double) * 3 * len, cudaMemcpyHostToDevice, s->stream));
			CK(cudaMemcpyAsync(s->stage3.p + 3 * ((size_t)nl + off), v + 3 * first, sizeof(double) * 3 * len, cudaMemcpyHostToDevice, s->stream));
		}
	} else {
		double *hx = s->mg_pinned, *hv = s->mg_pinned + 3 * (size_t)nl;
		for (int i = 0; i < nl; ++i) {
			const size_t g = 3 * (size_t)s->mg_local[i];
			hx[3 * i] = x[g]; hx[3 * i + 1] = x[g + 1]; hx[3 * i + 2] = x[g + 2];
			hv[3 * i] = v[g]; hv[3 * i + 1] = v[g + 1]; hv[3 * i + 2] = v[g + 2];
		}
		CK(cudaMemcpyAsync(s->stage3.p, s->mg_pinned, sizeof(double) * 6 * nl, cudaMemcpyHostToDevice, s->stream));
	}
	scatter3_to4_kernel<<<(nl + 255) / 256, 256, 0, s->stream>>>(nl, s->d_mg_local.p, s->stage3.p, s->x.p);
	scatter3_to4_kernel<<<(nl + 255) / 256, 256, 0, s->stream>>>(nl, s->d_mg_local.p, s->stage3.p + 3 * (size_t)nl, s->v.p);
	CK(cudaGetLastError());
	s->launches += 2;
}
void download_state_local(S *s, double *x, double *v)
{
	// owned AND ghost nodes come back: the next upload sends the ghosts up again, so the host must hold the values the
	// owners pushed into this rank's arrays at the end of the solve (they equal the owners' own values)
	const int no = (int)s->mg_local.size();
	if (!no) return;
	gather4_to3_kernel<<<(no + 255) / 256, 256, 0, s->stream>>>(no, s->d_mg_local.p, s->x.p, s->stage3.p);
	gather4_to3_kernel<<<(no + 255) / 256, 256, 0, s->stream>>>(no, s->d_mg_local.p, s->v.p, s->stage3.p + 3 * (size_t)no);
	CK(cudaGetLastError());
	s->launches += 2;
	if (s->mg_runs.size() <= 3 * 64) {
		for (size_t r = 0; r < s->mg_runs.size(); r += 3) {
			const size_t first = (size_t)s->mg_runs[r], len = (size_t)s->mg_runs[r + 1], off = (size_t)s->mg_runs[r + 2];
			CK(cudaMemcpyAsync(x + 3 * first, s->stage3.p + 3 * off, sizeof(double) * 3 * len, cudaMemcpyDeviceToHost, s->stream));
			CK(cudaMemcpyAsync(v + 3 * first, s->stage3.p + 3 * ((size_t)no + off), sizeof(double) * 3 * len, cudaMemcpyDeviceToHost, s->stream));
		}
		CK(cudaStreamSynchronize(s->stream));
		return;
	}
	CK(cudaMemcpyAsync(s->mg_pinned, s->stage3.p, sizeof(double) * 6 * no, cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	const double *hx = s->mg_pinned, *hv = s->mg_pinned + 3 * (size_t)no;
	for (int i = 0; i < no; ++i) {
		const size_t g = 3 * (size_t)s->mg_local[i];
		x[g] = hx[3 * i]; x[g + 1] = hx[3 * i + 1]; x[g + 2] = hx[3 * i + 2];
		v[g] = hv[3 * i]; v[g + 1] = hv[3 * i + 1]; v[g + 2] = hv[3 * i + 2];
	}
}

void upload_state(S *s, const double *x, const double *v)
{
	const int n = s->n_nodes;
	if (x) {
		CK(cudaMemcpyAsync(s->stage3.p, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream));
		pack3_to4_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->stage3.p, s->x.p);
		s->launches++;
	}
	if (v) {
		CK(cudaMemcpyAsync(s->stage3b.p, v, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream));
		pack3_to4_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->stage3b.p, s->v.p);
		s->launches++;
	}
	CK(cudaGetLastError());
}

void download_state(S *s, double *x, double *v)
{
	const int n = s->n_nodes;
	if (x) {
		unpack4_to3_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->x.p, s->stage3.p);
		CK(cudaMemcpyAsync(x, s->stage3.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s->stream));
		s->launches++;
	}
	if (v) {
		unpack4_to3_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->v.p, s->stage3b.p);
		CK(cudaMemcpyAsync(v, s->stage3b.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s->stream));
		s->launches++;
	}
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(s->stream));
}

template <typename E> void prox_tets_impl(S *s, int model, double mu, double lambda, double kappa, double bulk, int n, const double *z_in, double *z_out)
{
	int n_pad = pad32(n);
	std::vector<E> tmp((size_t)9 * n_pad, E(0));
	for (int e = 0; e < n; ++e) for (int k = 0; k < 9; ++k) tmp[(size_t)k * n_pad + e] = E(z_in[(size_t)9 * e + k]);
	DevBuf<E> d; d.upload(tmp, s->stream);
	Material<E> mat = Material<E>::make(mu, lambda, kappa, bulk);
	int threads = 128, blocks = (n + threads - 1) / threads;
	DevBuf<int> defer; defer.alloc((size_t)n + 1); defer.zero(s->stream);
#define ADMMB200_PROX_ONLY(M) \
	tet_prox_only_kernel<E, M><<<blocks, threads, 0, s->stream>>>(n, n_pad, d.p, mat, defer.p, defer.p + 1); \
	if (M != TET_LINEAR) tet_prox_only_deferred_kernel<E, M><<<std::min(blocks, 2 * s->n_sms), threads, 0, s->stream>>>(n_pad, d.p, mat, defer.p, defer.p + 1);
	switch (model) {
	case TET_LINEAR: ADMMB200_PROX_ONLY(TET_LINEAR) break;
	case TET_NEOHOOKEAN: ADMMB200_PROX_ONLY(TET_NEOHOOKEAN) break;
	case TET_STVK: ADMMB200_PROX_ONLY(TET_STVK) break;
	case TET_SPLINE_NH: ADMMB200_PROX_ONLY(TET_SPLINE_NH) break;
	case TET_SPLINE_STVK: ADMMB200_PROX_ONLY(TET_SPLINE_STVK) break;
	case TET_SPLINE_COROT: ADMMB200_PROX_ONLY(TET_SPLINE_COROT) break;
	default: throw std::runtime_error("unknown tet model");
	}
#undef ADMMB200_PROX_ONLY
	CK(cudaGetLastError());
	s->launches += 2;
	CK(cudaMemcpyAsync(tmp.data(), d.p, tmp.size() * sizeof(E), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	for (int e = 0; e < n; ++e) for (int k = 0; k < 9; ++k) z_out[(size_t)9 * e + k] = double(tmp[(size_t)k * n_pad + e]);
}

template <typename E> void prox_tris_impl(S *s, double lmin, double lmax, int n, const double *z_in, double *z_out)
{
	int n_pad = pad32(n);
	std::vector<E> tmp((size_t)6 * n_pad, E(0));
	for (int e = 0; e < n; ++e) for (int k = 0; k < 6; ++k) tmp[(size_t)k * n_pad + e] = E(z_in[(size_t)6 * e + k]);
	DevBuf<E> d; d.upload(tmp, s->stream);
	int threads = 128, blocks = (n + threads - 1) / threads;
	tri_prox_only_kernel<E><<<blocks, threads, 0, s->stream>>>(n, n_pad, d.p, E(lmin), E(lmax));
	CK(cudaGetLastError());
	s->launches++;
	CK(cudaMemcpyAsync(tmp.data(), d.p, tmp.size() * sizeof(E), cudaMemcpyDeviceToHost, s->stream));
	CK(cudaStreamSynchronize(s->stream));
	for (int e = 0; e < n; ++e) for (int k = 0; k < 6; ++k) z_out[(size_t)6 * e + k] = double(tmp[(size_t)k * n_pad + e]);
}

// copies element-precision SoA rows back into the reference's row layout
template <typename E> void gather_rows(S *s, bool want_z, double *out, long long n_out)
{
	auto fetch = [&](DevBuf<char> &buf, int k, int n, int n_pad, const std::vector<int> &row_off, const char *what) {
		require(!row_off.empty(), "debug_get z/u needs row_offset for every batch");
		require(buf.p != nullptr, what);
		std::vector<E> tmp((size_t)k * n_pad);
		CK(cudaMemcpy(tmp.data(), buf.p, tmp.size() * sizeof(E), cudaMemcpyDeviceToHost));
		for (int e = 0; e < n; ++e) for (int j = 0; j < k; ++j) {
			long long r = (long long)row_off[e] + j;
			require(r < n_out, "debug_get: output too small");
			out[r] = double(tmp[(size_t)j * n_pad + e]);
		}
	};
	for (auto t : s->tets) fetch(want_z ? t->d_z : t->d_u, 9, t->n, t->n_pad, t->row_off, "z not stored: call admm_b200_set_debug(1) before finalize");
	for (auto t : s->tris) fetch(want_z ? t->d_z : t->d_u, 6, t->n, t->n_pad, t->row_off, "z not stored: call admm_b200_set_debug(1) before finalize");
	if (s->pins.n) {
		require(!s->pins.row_off.empty(), "debug_get z/u needs row_offset for pins");
		std::vector<double> tmp((size_t)3 * s->pins.n);
		CK(cudaMemcpy(tmp.data(), want_z ? s->pins.d_z.p : s->pins.d_u.p, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
		for (int i = 0; i < s->pins.n; ++i) for (int j = 0; j < 3; ++j) {
			long long r = (long long)s->pins.row_off[i] + j;
			require(r < n_out, "debug_get: output too small");
			out[r] = tmp[(size_t)3 * i + j];
		}
	}
}

} // namespace

// ==========================================================================================
// C-ABI
// ==========================================================================================
extern "C" {

int admm_b200_version(void) { return 1; }

int admm_b200_create(int device, admm_b200_solver **out)
{
	if (!out) { g_create_error = "null out pointer"; return 1; }
	*out = nullptr;
	S *s = nullptr;
	int rc = guard(nullptr, [&]() {
		int count = 0;
		cudaError_t e = cudaGetDeviceCount(&count);
		if (e != cudaSuccess || count == 0) throw std::runtime_error(std::string("no CUDA device available (") + cudaGetErrorString(e) + "): the B200 path has no CPU fallback");
		if (device < 0 || device >= count) throw std::runtime_error("device index out of range");
		CK(cudaSetDevice(device));
		s = new S();
		s->device = device;
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, device));
		s->n_sms = prop.multiProcessorCount;
		s->gs_parts = s->n_sms;
		if (!prop.cooperativeLaunch) throw std::runtime_error("device lacks cooperative launch");
		CK(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
		s->stream = s->own_stream;
		s->barrier.alloc(4);
		s->iter_log.alloc(4096);
	});
	if (rc) { delete s; return rc; }
	*out = s;
	return 0;
}

void admm_b200_destroy(admm_b200_solver *s)
{
	if (!s) return;
	cudaSetDevice(s->device);
	cudaDeviceSynchronize();
	delete s;
}

const char *admm_b200_last_error(const admm_b200_solver *s) { return s ? s->err.c_str() : g_create_error.c_str(); }

int admm_b200_set_stream(admm_b200_solver *s, void *cuda_stream)
{
	return guard(s, [&]() {
		CK(cudaStreamSynchronize(s->stream));
		s->stream = cuda_stream ? (cudaStream_t)cuda_stream : s->own_stream;
	});
}

int admm_b200_synchronize(admm_b200_solver *s) { return guard(s, [&]() { CK(cudaStreamSynchronize(s->stream)); }); }

int admm_b200_set_nodes(admm_b200_solver *s, int n_nodes, const double *x, const double *v, const double *m)
{
	return guard(s, [&]() {
		require(!s->finalized, "set_nodes after finalize");
		require(n_nodes >= 1 && x && m, "**Solver Error: Problem with node data!");
		s->n_nodes = n_nodes;
		s->h_m.assign(m, m + (size_t)3 * n_nodes);
		s->h_x0.assign(x, x + (size_t)3 * n_nodes);
		s->x.alloc(n_nodes); s->v.alloc(n_nodes); s->cx.alloc(n_nodes); s->mxbar.alloc(n_nodes); s->b.alloc(n_nodes); s->m.alloc(n_nodes);
		s->stage3.alloc((size_t)6 * n_nodes); s->stage3b.alloc(std::max<size_t>((size_t)3 * n_nodes, 4096));
		s->v.zero(s->stream); s->b.zero(s->stream); s->mxbar.zero(s->stream);
		upload_state(s, x, v);
		CK(cudaStreamSynchronize(s->stream));
		CK(cudaMemcpyAsync(s->stage3.p, m, sizeof(double) * 3 * n_nodes, cudaMemcpyHostToDevice, s->stream));
		pack3_to4_kernel<<<(n_nodes + 255) / 256, 256, 0, s->stream>>>(n_nodes, s->stage3.p, s->m.p);
		CK(cudaMemcpyAsync(s->cx.p, s->x.p, sizeof(double4) * n_nodes, cudaMemcpyDeviceToDevice, s->stream));
		CK(cudaStreamSynchronize(s->stream));
	});
}

int admm_b200_add_tets(admm_b200_solver *s, int n, const int *idx, const double *dminv, const double *weight,
	int model, double mu, double lambda, double kappa, double bulk_modulus, const int *row_offset)
{
	return guard(s, [&]() {
		require(!s->finalized, "add_tets after finalize");
		require(n >= 0 && (n == 0 || (idx && dminv && weight)), "add_tets: null input");
		require(model >= 0 && model <= TET_SPLINE_COROT, "add_tets: unknown model");
		if (n == 0) return;
		for (int i = 0; i < 4 * n; ++i) require(idx[i] >= 0 && idx[i] < s->n_nodes, "add_tets: vertex index out of range (call set_nodes first)");
		for (int e = 0; e < n; ++e) require(weight[e] > 0.0, "**EnergyTerm::get_reduction Error: Some weight leq 0");
		TetBatchH *t = new TetBatchH();
		t->n = n; t->model = model; t->mu = mu; t->lambda = lambda; t->kappa = kappa; t->bulk = bulk_modulus;
		t->idx.assign(idx, idx + (size_t)4 * n);
		t->dminv.assign(dminv, dminv + (size_t)9 * n);
		t->w.assign(weight, weight + n);
		if (row_offset) t->row_off.assign(row_offset, row_offset + n);
		s->tets.push_back(t);
	});
}

int admm_b200_add_tris(admm_b200_solver *s, int n, const int *idx, const double *restpose, const double *weight,
	double limit_min, double limit_max, const int *row_offset)
{
	return guard(s, [&]() {
		require(!s->finalized, "add_tris after finalize");
		require(n >= 0 && (n == 0 || (idx && restpose && weight)), "add_tris: null input");
		require(!(limit_min > 1.0), "**TriEnergyTerm Error: Strain limit min should be -inf to 1");
		require(!(limit_max < 1.0), "**TriEnergyTerm Error: Strain limit max should be 1 to inf");
		if (n == 0) return;
		for (int i = 0; i < 3 * n; ++i) require(idx[i] >= 0 && idx[i] < s->n_nodes, "add_tris: vertex index out of range (call set_nodes first)");
		for (int e = 0; e < n; ++e) require(weight[e] > 0.0, "**EnergyTerm::get_reduction Error: Some weight leq 0");
		TriBatchH *t = new TriBatchH();
		t->n = n; t->limit_min = limit_min; t->limit_max = limit_max;
		t->idx.assign(idx, idx + (size_t)3 * n);
		t->rest.assign(restpose, restpose + (size_t)4 * n);
		t->w.assign(weight, weight + n);
		if (row_offset) t->row_off.assign(row_offset, row_offset + n);
		s->tris.push_back(t);
	});
}

int admm_b200_add_pins(admm_b200_solver *s, int n, const int *idx, const double *pos, const double *weight, const int *row_offset)
{
	return guard(s, [&]() {
		require(!s->finalized, "add_pins after finalize");
		require(n >= 0 && (n == 0 || (idx && pos && weight)), "add_pins: null input");
		for (int i = 0; i < n; ++i) require(idx[i] >= 0 && idx[i] < s->n_nodes, "add_pins: vertex index out of range");
		PinsH &p = s->pins;
		p.idx.insert(p.idx.end(), idx, idx + n);
		p.pos.insert(p.pos.end(), pos, pos + (size_t)3 * n);
		p.w.insert(p.w.end(), weight, weight + n);
		p.active.insert(p.active.end(), n, (unsigned char)1);
		if (row_offset) p.row_off.insert(p.row_off.end(), row_offset, row_offset + n);
		p.n += n;
	});
}

int admm_b200_update_pins(admm_b200_solver *s, int n, const double *pos, const unsigned char *active)
{
	return guard(s, [&]() {
		PinsH &p = s->pins;
		require(n == p.n, "update_pins: pin count differs from add_pins");
		if (pos) p.pos.assign(pos, pos + (size_t)3 * n);
		if (active) p.active.assign(active, active + n);
		if (s->finalized && n) {
			// synchronous so the borrowed host arrays can be released on return
			CK(cudaMemcpyAsync(p.d_pos.p, p.pos.data(), sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream));
			CK(cudaMemcpyAsync(p.d_active.p, p.active.data(), n, cudaMemcpyHostToDevice, s->stream));
			CK(cudaStreamSynchronize(s->stream));
		}
	});
}

int admm_b200_set_gs_pins(admm_b200_solver *s, int n, const int *idx, const double *pos)
{
	return guard(s, [&]() {
		require(n >= 0 && (n == 0 || (idx && pos)), "set_gs_pins: null input");
		s->gs_pin_idx.assign(idx, idx + n);
		s->gs_pin_pos.assign(pos, pos + (size_t)3 * n);
		if (s->finalized && s->linsolver == ADMM_B200_MCGS) upload_gs_pins(s);
	});
}

int admm_b200_set_surface_inds(admm_b200_solver *s, int n, const int *idx)
{
	return guard(s, [&]() {
		require(!s->finalized, "set_surface_inds after finalize");
		require(n >= 0 && (n == 0 || idx), "set_surface_inds: null input");
		s->surface_inds.assign(idx, idx + n);
	});
}

int admm_b200_set_constraint_weight(admm_b200_solver *s, double constraint_w)
{
	return guard(s, [&]() {
		require(!s->finalized, "set_constraint_weight after finalize");
		s->constraint_w = constraint_w;
	});
}

int admm_b200_add_obstacle(admm_b200_solver *s, int kind, const double *params)
{
	return guard(s, [&]() {
		require(kind == ADMM_B200_FLOOR || kind == ADMM_B200_SPHERE, "add_obstacle: unknown kind");
		require((int)s->obstacles.size() < ADMMB200_MAX_OBSTACLES, "add_obstacle: too many obstacles");
		Obstacle o; o.kind = kind;
		for (int i = 0; i < 4; ++i) o.p[i] = (kind == ADMM_B200_FLOOR && i > 0) ? 0.0 : params[i];
		s->obstacles.push_back(o);
	});
}

int admm_b200_add_wind(admm_b200_solver *s, const int *tris, int n_tris, const double *direction, int *id)
{
	return guard(s, [&]() {
		require(n_tris >= 0 && (tris || n_tris == 0) && direction, "add_wind: bad arguments");
		require(s->world == 1, "add_wind: explicit forces are single-GPU");
		std::unique_ptr<S::Wind> w(new S::Wind());
		w->tris.assign(tris, tris + 3 * (size_t)n_tris);
		for (int a = 0; a < 3; ++a) w->dir[a] = direction[a];
		s->winds.push_back(std::move(w));
		if (id) *id = (int)s->winds.size() - 1;
	});
}

int admm_b200_set_wind_direction(admm_b200_solver *s, int id, const double *direction)
{
	return guard(s, [&]() {
		require(id >= 0 && id < (int)s->winds.size() && direction, "set_wind_direction: no such wind force");
		for (int a = 0; a < 3; ++a) s->winds[id]->dir[a] = direction[a];
	});
}

int admm_b200_set_system(admm_b200_solver *s, int n, const int *rowptr, const int *cols, const double *vals)
{
	return guard(s, [&]() {
		require(!s->finalized, "set_system after finalize");
		require(n > 0 && rowptr && cols && vals, "**NodalMultiColorGS Error: Bad dimensions in A");
		s->sys_n = n;
		s->L_rowptr.assign(rowptr, rowptr + n + 1);
		s->L_cols.assign(cols, cols + rowptr[n]);
		s->L_vals.assign(vals, vals + rowptr[n]);
	});
}

int admm_b200_set_colors(admm_b200_solver *s, int n_colors, const int *offsets, const int *nodes)
{
	return guard(s, [&]() {
		require(!s->finalized, "set_colors after finalize");
		require(n_colors > 0 && offsets && nodes, "set_colors: null input");
		s->n_colors = n_colors;
		s->color_off.assign(offsets, offsets + n_colors + 1);
		s->color_nodes.assign(nodes, nodes + offsets[n_colors]);
	});
}

int admm_b200_set_ldlt(admm_b200_solver *s, int n, const int *perm, const int *Lp, const int *Li, const double *Lx, const double *D)
{
	return guard(s, [&]() {
		require(!s->finalized, "set_ldlt after finalize");
		require(n > 0 && perm && Lp && D, "**LDLTSolver Error: Bad dimensions in A");
		s->ld_n = n;
		s->ld_perm.assign(perm, perm + n);
		s->ld_Lp.assign(Lp, Lp + n + 1);
		s->ld_Li.assign(Li, Li + Lp[n]);
		s->ld_Lx.assign(Lx, Lx + Lp[n]);
		s->ld_D.assign(D, D + n);
		for (int i = 0; i < n; ++i) require(D[i] != 0.0, "set_ldlt: zero pivot");
		s->have_ldlt = true;
	});
}

int admm_b200_set_rank(admm_b200_solver *s, int rank, int world)
{
	return guard(s, [&]() {
		require(!s->finalized, "set_rank after finalize");
		require(world >= 1 && world <= ADMMB200_MAX_RANKS && rank >= 0 && rank < world, "set_rank: rank/world out of range (at most 8 ranks)");
		s->rank = rank; s->world = world;
	});
}

int admm_b200_device_sms(const admm_b200_solver *s) { return s ? s->n_sms : 0; }

int admm_b200_mgpu_nodes(const admm_b200_solver *s, int *n_owned, int *n_ghost)
{
	if (!s) return 1;
	if (n_owned) *n_owned = s->world > 1 ? (int)s->mg_owned.size() : s->n_nodes;
	if (n_ghost) *n_ghost = s->world > 1 ? (int)(s->mg_local.size() - s->mg_owned.size()) : 0;
	return 0;
}

int admm_b200_gs_parts(const admm_b200_solver *s) { return s ? s->gs_parts : 0; }

int admm_b200_set_gs_parts(admm_b200_solver *s, int n_parts)
{
	return guard(s, [&]() {
		require(!s->finalized, "set_gs_parts after finalize");
		require(n_parts >= 0 && n_parts <= s->n_sms, "set_gs_parts: between 1 and the SM count (0 = one part per SM)");
		s->gs_parts = n_parts > 0 ? n_parts : s->n_sms;
	});
}

namespace {
struct IpcBlob { cudaIpcMemHandle_t x, dglob, flags; int rank, n_nodes; };
static_assert(sizeof(IpcBlob) <= ADMM_B200_IPC_BYTES, "IPC blob too large");
}

int admm_b200_mgpu_export(admm_b200_solver *s, void *blob)
{
	return guard(s, [&]() {
		require(s->finalized && s->world > 1 && blob, "mgpu_export: needs a finalized multi-rank solver");
		require(s->gs_resident && s->res_dglob.p && s->mg_flags.p, "mgpu_export: the resident fp32 Gauss-Seidel is not active");
		IpcBlob b;
		std::memset(&b, 0, sizeof(b));
		CK(cudaIpcGetMemHandle(&b.x, s->cx.p));
		CK(cudaIpcGetMemHandle(&b.dglob, s->res_dglob.p));
		CK(cudaIpcGetMemHandle(&b.flags, s->mg_flags.p));
		b.rank = s->rank; b.n_nodes = s->n_nodes;
		std::memset(blob, 0, ADMM_B200_IPC_BYTES);
		std::memcpy(blob, &b, sizeof(b));
	});
}

int admm_b200_mgpu_import(admm_b200_solver *s, int peer_rank, const void *blob)
{
	return guard(s, [&]() {
		require(s->finalized && s->world > 1 && blob, "mgpu_import: needs a finalized multi-rank solver");
		require(peer_rank >= 0 && peer_rank < s->world, "mgpu_import: peer rank out of range");
		if (peer_rank == s->rank) return;
		IpcBlob b;
		std::memcpy(&b, blob, sizeof(b));
		require(b.rank == peer_rank && b.n_nodes == s->n_nodes, "mgpu_import: blob does not belong to that rank / mesh");
		void *px = nullptr, *pd = nullptr, *pf = nullptr;
		CK(cudaIpcOpenMemHandle(&px, b.x, cudaIpcMemLazyEnablePeerAccess)); s->ipc_opened.push_back(px);
		CK(cudaIpcOpenMemHandle(&pd, b.dglob, cudaIpcMemLazyEnablePeerAccess)); s->ipc_opened.push_back(pd);
		CK(cudaIpcOpenMemHandle(&pf, b.flags, cudaIpcMemLazyEnablePeerAccess)); s->ipc_opened.push_back(pf);
		s->peer_x[peer_rank] = (double4 *)px; s->peer_dglob[peer_rank] = (uint2 *)pd; s->peer_flags[peer_rank] = (unsigned int *)pf;
	});
}

int admm_b200_mgpu_ready(admm_b200_solver *s)
{
	return guard(s, [&]() {
		require(s->finalized && s->world > 1, "mgpu_ready: needs a finalized multi-rank solver");
		for (int q = 0; q < s->world; ++q) if (q != s->rank) require(s->peer_x[q] && s->peer_dglob[q] && s->peer_flags[q], "mgpu_ready: a peer has not been imported");
		s->mg_ready = true;
	});
}

// Host-only: node -> part of the resident plan (csrc/partition.hpp) for n_parts parts.  With n_parts =
// world * admm_b200_device_sms(), part / sms is the rank that owns the node.
int admm_b200_plan_parts(int n, const int *rowptr, const int *cols, const double *vals, const double *pos3, int n_parts, int *part_of)
{
	try {
		if (n <= 0 || !rowptr || !cols || !vals || !pos3 || !part_of || n_parts <= 0) throw std::runtime_error("plan_parts: bad arguments");
		// EXACTLY the partition plan_resident computes at finalize: same code, same lane count
		std::vector<int> part = plan_partition(n, rowptr, cols, vals, pos3, n_parts, plan_default_lanes());
		std::copy(part.begin(), part.end(), part_of);
		return 0;
	} catch (std::exception &e) { g_create_error = e.what(); return 1; }
}

// Host-only: what rank `rank` of `world` would exchange.  mask_out[g] = ranks (bit q) that read node g, for
// nodes this rank owns; ghost_out[g] = 1 for nodes this rank reads from another rank; owner_out[g] = rank.
int admm_b200_mgpu_plan_check(int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off, const int *color_nodes,
	const double *pos3, int sms, int world, int rank, unsigned int *mask_out, int *ghost_out, int *owner_out)
{
	try {
		ResidentPlan R = plan_resident(n, rowptr, cols, vals, n_colors, color_off, color_nodes, pos3, sms * world, plan_default_lanes());
		std::vector<unsigned int> m = dest_masks(R, n, sms, rank);
		for (int i = 0; i < n; ++i) { mask_out[i] = m[i]; ghost_out[i] = 0; owner_out[i] = R.part_of[i] / sms; }
		for (int p = rank * sms; p < (rank + 1) * sms; ++p) {
			const PartDesc &d = R.parts[p];
			for (int h = 0; h < d.n_halo; ++h) { int g = R.gid[d.gid_off + d.n_own + h]; if (R.part_of[g] / sms != rank) ghost_out[g] = 1; }
		}
		return 0;
	} catch (std::exception &e) { g_create_error = e.what(); return 1; }
}

int admm_b200_set_debug(admm_b200_solver *s, int store_z)
{
	return guard(s, [&]() { require(!s->finalized, "set_debug after finalize"); s->store_z = store_z != 0; });
}

int admm_b200_finalize(admm_b200_solver *s, double dt, int linsolver, int gs_iters, double gs_omega, double gs_tol, int precision)
{
	return guard(s, [&]() {
		require(!s->finalized, "finalize called twice");
		require(s->n_nodes > 0, "**Solver Error: Problem with node data!");
		require(linsolver >= 0 && linsolver <= 2, "unknown linsolver");
		require(precision == ADMM_B200_FP32 || precision == ADMM_B200_FP64, "unknown precision");
		if (dt <= 0.0) dt = 1.0 / 24.0; // src/Solver.cpp:175-179
		s->dt = dt; s->linsolver = linsolver; s->gs_iters = gs_iters; s->gs_omega = gs_omega; s->gs_tol = gs_tol; s->precision = precision;
		if (const char *e = getenv("ADMM_B200_TET_MINBLOCKS")) s->tet_minblocks = atoi(e);
		// No collisions with the LDLT solver (src/Solver.cpp:249-254)
		if (linsolver == ADMM_B200_LDLT) require(s->obstacles.empty(), "**Solver::add_obstacle Error: No collisions with LDLT solver");
		build_incidence(s);
		if (precision == ADMM_B200_FP64) upload_elements<double>(s); else upload_elements<float>(s);
		upload_pins(s);
		if (linsolver == ADMM_B200_MCGS) build_mcgs(s); else build_ldlt(s);
		CK(cudaStreamSynchronize(s->stream));
		s->finalized = true;
	});
}

int admm_b200_step(admm_b200_solver *s, int admm_iters, double gravity, admm_b200_runtime *runtime)
{
	return guard(s, [&]() { do_step(s, admm_iters, gravity, runtime); });
}

int admm_b200_upload_state(admm_b200_solver *s, const double *x, const double *v)
{
	return guard(s, [&]() {
		require(s->n_nodes > 0, "no nodes");
		if (s->world > 1 && s->finalized && x && v && !s->mg_local.empty()) upload_state_local(s, x, v); else upload_state(s, x, v);
		CK(cudaStreamSynchronize(s->stream));
	});
}

int admm_b200_download_state(admm_b200_solver *s, double *x, double *v)
{
	return guard(s, [&]() {
		require(s->n_nodes > 0, "no nodes");
		// several ranks: only the nodes this rank owns are written (the rest of the device arrays carries no elastic forces)
		if (s->world > 1 && s->finalized && x && v && !s->mg_local.empty()) download_state_local(s, x, v); else download_state(s, x, v);
	});
}

int admm_b200_pin_host(admm_b200_solver *s, void *ptr, unsigned long long bytes)
{
	return guard(s, [&]() {
		require(ptr && bytes > 0, "pin_host: null range");
		CK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
	});
}

int admm_b200_unpin_host(admm_b200_solver *s, void *ptr)
{
	return guard(s, [&]() { require(ptr != nullptr, "unpin_host: null"); CK(cudaHostUnregister(ptr)); });
}

int admm_b200_step_host(admm_b200_solver *s, int admm_iters, double gravity, double *x, double *v, admm_b200_runtime *runtime)
{
	return guard(s, [&]() {
		require(x && v, "step_host: null state");
		if (s->world > 1 && !s->mg_local.empty()) {
			upload_state_local(s, x, v);
			do_step(s, admm_iters, gravity, runtime);
			download_state_local(s, x, v);
			return;
		}
		upload_state(s, x, v);
		do_step(s, admm_iters, gravity, runtime);
		download_state(s, x, v);
	});
}

int admm_b200_prox_tets(admm_b200_solver *s, int model, double mu, double lambda, double kappa, double bulk_modulus, int precision, int n, const double *z_in, double *z_out)
{
	return guard(s, [&]() {
		require(n >= 0 && (n == 0 || (z_in && z_out)), "prox_tets: null input");
		if (n == 0) return;
		if (precision == ADMM_B200_FP64) prox_tets_impl<double>(s, model, mu, lambda, kappa, bulk_modulus, n, z_in, z_out);
		else prox_tets_impl<float>(s, model, mu, lambda, kappa, bulk_modulus, n, z_in, z_out);
	});
}

int admm_b200_prox_tris(admm_b200_solver *s, double limit_min, double limit_max, int precision, int n, const double *z_in, double *z_out)
{
	return guard(s, [&]() {
		require(n >= 0 && (n == 0 || (z_in && z_out)), "prox_tris: null input");
		if (n == 0) return;
		if (precision == ADMM_B200_FP64) prox_tris_impl<double>(s, limit_min, limit_max, n, z_in, z_out);
		else prox_tris_impl<float>(s, limit_min, limit_max, n, z_in, z_out);
	});
}

int admm_b200_linsolve(admm_b200_solver *s, double *x, const double *b, int *iters)
{
	return guard(s, [&]() {
		require(s->finalized, "linsolve before finalize");
		const int n = s->n_nodes;
		CK(cudaMemcpyAsync(s->stage3.p, x, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream));
		pack3_to4_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->stage3.p, s->cx.p);
		CK(cudaMemcpyAsync(s->stage3b.p, b, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream));
		pack3_to4_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->stage3b.p, s->b.p);
		launch_global(s);
		unpack4_to3_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(n, s->cx.p, s->stage3.p);
		CK(cudaMemcpyAsync(x, s->stage3.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, s->stream));
		int it = 1;
		if (s->linsolver == ADMM_B200_MCGS || (s->linsolver == ADMM_B200_UZAWA && !s->obstacles.empty())) CK(cudaMemcpyAsync(&it, s->gs_iters_done.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
		CK(cudaStreamSynchronize(s->stream));
		if (iters) *iters = it;
	});
}

int admm_b200_debug_get(admm_b200_solver *s, const char *name, double *out, long long n_out)
{
	return guard(s, [&]() {
		require(name && out, "debug_get: null input");
		CK(cudaStreamSynchronize(s->stream));
		std::string nm(name);
		const int n = s->n_nodes;
		if (nm == "z" || nm == "u") {
			require(s->finalized, "debug_get before finalize");
			if (s->precision == ADMM_B200_FP64) gather_rows<double>(s, nm == "z", out, n_out); else gather_rows<float>(s, nm == "z", out, n_out);
			return;
		}
		if (nm == "ldlt_prof") {
			require(s->lb_prof.p != nullptr, "debug_get ldlt_prof: set ADMM_B200_LDLT_PROF=1 before finalize");
			require(n_out >= (long long)s->lb_prof.n, "debug_get: output too small");
			std::vector<unsigned long long> tmp(s->lb_prof.n);
			CK(cudaMemcpy(tmp.data(), s->lb_prof.p, tmp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
			const unsigned long long t0 = tmp[0];
			for (size_t i = 0; i < tmp.size(); ++i) out[i] = tmp[i] ? (double)(tmp[i] - t0) : -1.0; // ns since the start of the solve
			return;
		}
		if (nm == "gs_prof") {
			require(s->res_prof.p != nullptr, "debug_get gs_prof: set ADMM_B200_GS_PROF=1 before finalize");
			require(n_out >= (long long)s->res_prof.n, "debug_get: output too small");
			std::vector<unsigned long long> tmp(s->res_prof.n);
			CK(cudaMemcpy(tmp.data(), s->res_prof.p, tmp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
			for (size_t i = 0; i < tmp.size(); ++i) out[i] = (double)tmp[i];
			return;
		}
		const double4 *src = nullptr;
		if (nm == "b") src = s->b.p; else if (nm == "x") src = s->cx.p; else if (nm == "v") src = s->v.p; else if (nm == "x0") src = s->x.p;
		else throw std::runtime_error("debug_get: unknown array name");
		require(n_out >= 3LL * n, "debug_get: output too small");
		std::vector<double4> tmp(n);
		CK(cudaMemcpy(tmp.data(), src, sizeof(double4) * n, cudaMemcpyDeviceToHost));
		for (int i = 0; i < n; ++i) { out[3 * i] = tmp[i].x; out[3 * i + 1] = tmp[i].y; out[3 * i + 2] = tmp[i].z; }
	});
}

int admm_b200_time_kernels(admm_b200_solver *s, int reps, double *out_ms)
{
	return guard(s, [&]() {
		require(s->finalized && reps > 0 && out_ms, "time_kernels: bad arguments");
		cudaEvent_t e0 = get_event(s, 0), e1 = get_event(s, 1);
		float ms;
		CK(cudaStreamSynchronize(s->stream));
		CK(cudaEventRecord(e0, s->stream));
		for (int r = 0; r < reps; ++r) launch_local(s);
		CK(cudaEventRecord(e1, s->stream));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventElapsedTime(&ms, e0, e1)); out_ms[0] = ms / reps;
		CK(cudaEventRecord(e0, s->stream));
		for (int r = 0; r < reps; ++r) launch_assemble(s);
		CK(cudaEventRecord(e1, s->stream));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventElapsedTime(&ms, e0, e1)); out_ms[1] = ms / reps;
		CK(cudaEventRecord(e0, s->stream));
		for (int r = 0; r < reps; ++r) launch_global(s);
		CK(cudaEventRecord(e1, s->stream));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventElapsedTime(&ms, e0, e1)); out_ms[2] = ms / reps;
	});
}

long long admm_b200_launch_count(const admm_b200_solver *s) { return s ? s->launches : 0; }

int admm_b200_set_deferred_timers(admm_b200_solver *s, int on)
{
	return guard(s, [&]() {
		if (!s->tblocks.empty()) { admm_b200_runtime tmp; sum_timer_blocks(s, &tmp); }
		s->deferred_timers = on != 0;
		s->deferred_stride = on > 1 ? on : 1; s->deferred_count = 0;
		s->pend = admm_b200_runtime{0, 0, 0, 0, 0, 0}; s->pend_steps = 0;
		for (int k = 0; k < 3; ++k) { s->pend_kernel_ms[k] = 0; s->pend_kernel_n[k] = 0; }
	});
}

int admm_b200_collect_timers(admm_b200_solver *s, admm_b200_runtime *sum, int *steps)
{
	return guard(s, [&]() {
		require(sum != nullptr, "collect_timers: null output");
		sum_timer_blocks(s, sum);
		sum->global_ms += s->pend.global_ms; sum->local_ms += s->pend.local_ms; sum->assemble_ms += s->pend.assemble_ms; sum->step_ms += s->pend.step_ms; sum->inner_iters += s->pend.inner_iters;
		for (int k = 0; k < 3; ++k) { s->kernel_ms[k] += s->pend_kernel_ms[k]; s->kernel_n[k] += s->pend_kernel_n[k]; s->pend_kernel_ms[k] = 0; s->pend_kernel_n[k] = 0; }
		if (steps) *steps = s->timer_steps + s->pend_steps;
		s->pend = admm_b200_runtime{0, 0, 0, 0, 0, 0}; s->pend_steps = 0;
	});
}

int admm_b200_kernel_times(admm_b200_solver *s, double *out_ms, long long *out_n)
{
	return guard(s, [&]() {
		require(out_ms && out_n, "kernel_times: bad arguments");
		for (int k = 0; k < 3; ++k) { out_ms[k] = s->kernel_ms[k]; out_n[k] = s->kernel_n[k]; }
	});
}

// Host-only check of the resident plan (no device): builds the plan, walks it exactly like
// mcgs_resident_kernel's gather and returns max |(L_offdiag x)_plan - (L_offdiag x)_csr| over all nodes
// for the given x (n values); stats = {shared bytes needed, max own, max halo, entries, nnz, parts used}.
// Host-only model of the barrier-free schedule (csrc/dataflow_plan.hpp): plans it for the given scalar matrix and
// colouring, executes `sweeps` SOR sweeps under `n_seeds` random legal schedules and requires bit-identical results to
// colour-by-colour sweeps.  stats = {max own-slice dependencies, max halo references per slice, tasks executed, parts used}.
int admm_b200_dataflow_check(int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off, const int *color_nodes,
	const double *pos3, int n_parts, int n_warps, int sweeps, int n_seeds, long long *stats)
{
	try {
		ResidentPlan R = plan_resident(n, rowptr, cols, vals, n_colors, color_off, color_nodes, pos3, n_parts, 1);
		DataflowPlan D = plan_dataflow(R, n_colors, n_warps);
		std::vector<double> diag(n, 1.0), rhs(n);
		for (int i = 0; i < n; ++i) for (int q = rowptr[i]; q < rowptr[i + 1]; ++q) if (cols[q] == i) diag[i] = vals[q];
		long long tasks = 0, used = 0;
		for (const PartDesc &d : R.parts) if (d.n_own) ++used;
		for (int sd = 0; sd < n_seeds; ++sd) {
			std::mt19937 rng(1000 + sd);
			std::uniform_real_distribution<double> U(-1.0, 1.0);
			for (int i = 0; i < n; ++i) rhs[i] = U(rng);
			tasks += simulate_dataflow(R, D, n, n_colors, diag.data(), rhs.data(), 1.9, sweeps, (unsigned int)(77 + sd), nullptr);
		}
		if (stats) { stats[0] = (long long)D.max_deps; stats[1] = (long long)D.max_halo_refs; stats[2] = tasks; stats[3] = used; }
		return 0;
	} catch (std::exception &e) {
		g_create_error = e.what();
		return 1;
	}
}

int admm_b200_plan_check(int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off, const int *color_nodes,
	const double *pos3, int n_parts, int val_bytes, int lanes, const double *x, double *max_err, long long *stats, int *part_of)
{
	try {
		if (lanes != 1 && lanes != 2 && lanes != 4) throw std::runtime_error("plan_check: lanes must be 1, 2 or 4");
		ResidentPlan R = plan_resident(n, rowptr, cols, vals, n_colors, color_off, color_nodes, pos3, n_parts, lanes);
		const int T = lanes, G = 32 / T;
		double worst = 0;
		std::vector<char> seen(n, 0);
		long long used = 0;
		for (const PartDesc &d : R.parts) {
			if (d.n_own) ++used;
			const int *gid = R.gid.data() + d.gid_off;
			const int *srow = R.slice_row.data() + d.slice_off;
			const short *snode = R.slice_node.data() + d.snode_off;
			const int *cs = R.color_slice.data() + d.cslice_off;
			if (cs[2 * n_colors] != d.n_slices) throw std::runtime_error("plan: colour table does not cover all slices");
			for (int c = 0; c < 2 * n_colors; ++c) for (int sl = cs[c]; sl < cs[c + 1]; ++sl) for (int g = 0; g < G; ++g) {
				int l = snode[sl * G + g];
				if (l < 0) continue;
				int node = gid[l];
				if (seen[node]) throw std::runtime_error("plan: node updated twice");
				seen[node] = 1;
				double acc = 0;
				if ((c % 2 == 0) != (R.part_of[node] == R.part_of[node] && [&]() { for (int q = rowptr[node]; q < rowptr[node + 1]; ++q) if (cols[q] != node && vals[q] != 0.0 && R.part_of[cols[q]] != R.part_of[node]) return false; return true; }()))
					throw std::runtime_error("plan: interior/boundary classification is wrong");
				for (int r = srow[sl]; r < srow[sl + 1]; ++r) for (int t = 0; t < T; ++t) {
					size_t e = (size_t)d.ent_off + (size_t)r * 32 + g * T + t;
					int cl = R.col[e];
					if (cl >= d.n_own + d.n_halo) throw std::runtime_error("plan: column out of range");
					double v = val_bytes == 4 ? (double)(float)R.val[e] : R.val[e];
					acc += v * x[gid[cl]];
				}
				double ref = 0;
				for (int q = rowptr[node]; q < rowptr[node + 1]; ++q) if (cols[q] != node) ref += (val_bytes == 4 ? (double)(float)vals[q] : vals[q]) * x[cols[q]];
				worst = std::max(worst, std::abs(acc - ref));
			}
		}
		for (int i = 0; i < n; ++i) if (!seen[i]) throw std::runtime_error("plan: a node is never updated");
		if (max_err) *max_err = worst;
		if (stats) { stats[0] = (long long)R.smem_bytes(n_colors, val_bytes); stats[1] = (long long)R.max_own; stats[2] = (long long)R.max_halo; stats[3] = (long long)R.entries; stats[4] = (long long)R.nnz; stats[5] = used; }
		if (part_of) std::copy(R.part_of.begin(), R.part_of.end(), part_of);
		return 0;
	} catch (std::exception &e) {
		g_create_error = e.what();
		return 1;
	}
}

// Host-only: modelled shared-memory cycles of the float4 gathers of one sweep (quarter-warp phases x bank-group
// multiplicity, partition.hpp: rowstep_cycles) with the entries in matrix order and after the conflict-aware scheduling,
// plus the conflict-free minimum (4 per row-step).  out[8] = {before, after, minimum, tiled plan: shared bytes, max slices, ELL fill x 1000, owned plan: shared bytes, tiled plan: max rows}.
int admm_b200_plan_bank_stats(int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off, const int *color_nodes,
	const double *pos3, int n_parts, long long *out)
{
	try {
		ResidentPlan R = plan_resident(n, rowptr, cols, vals, n_colors, color_off, color_nodes, pos3, n_parts, 1);
		out[0] = R.cycles_before; out[1] = R.cycles_after; out[2] = 4 * (long long)(R.entries / 32);
		// the tiled kernel's plan (4 lanes per node): shared memory per part, slices, ELL fill in 1/1000
		ResidentPlan R4 = plan_resident(n, rowptr, cols, vals, n_colors, color_off, color_nodes, pos3, n_parts, 4);
		out[3] = (long long)R4.smem_bytes(n_colors, 4, 2); out[4] = (long long)R4.max_slices; out[5] = R4.entries ? (long long)(1000.0 * (double)R4.nnz / (double)R4.entries) : 0;
		out[6] = (long long)R.smem_bytes(n_colors, 4, 1); out[7] = (long long)R4.max_rows;
		{ long long c = 0; for (size_t r = 0; r < R4.entries / 32; ++r) c += detail::rowstep_cycles(&R4.col[r * 32]); out[8] = c; out[9] = 4 * (long long)(R4.entries / 32); }
		return 0;
	} catch (std::exception &e) {
		g_create_error = e.what();
		return 1;
	}
}

// Host-only: plans the block solve for a factor (ldlt_blocks.hpp) and applies it on the host to one right-hand side
// (n values), exactly as the device kernel walks it.  stats[8] = {blocks, largest block, forward levels, backward levels,
// entries outside the diagonal blocks, entries of the inverted diagonal blocks, cut level of the bottom forest, segments}.
int admm_b200_ldlt_blocks_check(int n, const int *perm, const int *Lp, const int *Li, const double *Lx, const double *D, const double *b, double *x, long long *stats)
{
	try {
		LdltBlockPlan B = plan_ldlt_blocks(n, Lp, Li, Lx);
		ldlt_blocks_solve_host(B, perm, D, b, x);
		if (stats) { stats[0] = B.n_blocks; stats[1] = B.max_block; stats[2] = B.n_levels_f; stats[3] = B.n_levels_b; stats[4] = B.nnz_out; stats[5] = B.nnz_inv; stats[6] = B.cut; stats[7] = (long long)B.seg_ptr.back(); }
		return 0;
	} catch (std::exception &e) {
		g_create_error = e.what();
		return 1;
	}
}

const char *admm_b200_solver_info(const admm_b200_solver *s) { return s ? s->gs_info.c_str() : ""; }

} // extern "C"
