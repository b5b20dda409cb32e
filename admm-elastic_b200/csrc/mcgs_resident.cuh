// mcgs_resident.cuh -- NodalMultiColorGS::solve (src/NodalMultiColorGS.hpp:60-146) with the matrix and
// the iterate resident in shared memory.
//
// One persistent CTA per SM owns one part of the mesh (partition.hpp).  At launch it pulls the rows of
// its nodes (16-bit local column + value, sliced ELL) into shared memory with TMA bulk copies
// (cp.async.bulk, one mbarrier) and the positions of its nodes with plain loads; after that every
// sweep x colour pass reads matrix and neighbours from shared memory (LDS ~30 cycles instead of an L2
// round trip of several hundred) and goes to L2 only for halo neighbours, b and the result.  Updated
// positions are written both to shared memory (this part's later colours) and to global memory (other
// parts' halo reads, and the caller).  A grid barrier separates colours, as in mcgs_kernel.
//
// Semantics are those of mcgs_kernel (kernels.cuh): pins override, passive obstacles inside the sweep,
// SOR with omega, the reference's convergence test evaluated lazily with an identical outcome.
#pragma once
#include "kernels.cuh"
#include "partition.hpp"
#include <cstdint>

namespace admmb200 {

#define ADMMB200_RES_THREADS 768

struct McgsResParams {
	McgsParams base;            // x, b, diag, pins, obstacles, barrier, residual slots, omega, iters, tol2, n_colors
	const PartDesc *parts;      // [gridDim.x]
	const uint16_t *col;        // all parts, 32-entry rows
	const void *val;            // float or double, same indexing as col
	const int *gid, *slice_row, *color_slice, *nbr;
	const short *slice_node;
	unsigned int *part_epoch;   // [gridDim.x * 8] one flag per part (32-byte stride), zeroed before launch
	unsigned int *sweep_flag;   // [iters] set to 1 by any part that proves "not converged yet" for that sweep
	unsigned int *sweep_arrive; // [iters] parts that have finished that sweep
	unsigned long long *prof;   // NULL, or [gridDim.x * 16] clock cycles of thread 0: waiting, boundary compute, publishing, total, interior compute
};

// Point-to-point ordering between neighbouring parts, replacing a grid barrier per colour pass.
// publish: this part has finished pass `epoch` (all its reads of neighbours' values and all its writes).
// wait:    every neighbour has finished pass `epoch`, so (a) their values of that pass are visible and
//          (b) they no longer read the values this part is about to overwrite.
// Only the boundary warps (warps [0, n_bwarps)) wait: they synchronise among themselves with named
// barrier 1, so the interior warps keep computing while the flags are in flight.
__device__ __forceinline__ void named_sync(int id, int n_threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory"); }

__device__ __forceinline__ void part_publish(unsigned int *part_epoch, unsigned int epoch)
{
	// caller has just passed a __syncthreads(): every warp's writes of this pass are done
	if (threadIdx.x == 0) { fence_acq_rel_gpu(); st_relaxed_u32(part_epoch + 8 * blockIdx.x, epoch); }
}
__device__ __forceinline__ void part_wait(const unsigned int *part_epoch, const int *s_nbr, int n_nbr, unsigned int epoch, int n_bthreads)
{
	if ((int)threadIdx.x < n_nbr) {
		for (int i = threadIdx.x; i < n_nbr; i += n_bthreads) {
			const unsigned int *f = part_epoch + 8 * s_nbr[i];
			while (ld_relaxed_u32(f) < epoch) { }
		}
		fence_acq_rel_gpu();
	}
	named_sync(1, n_bthreads);
}

// ---- TMA bulk copy (cp.async.bulk) + mbarrier, used once per launch to stage the matrix ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE;\n"
		"bra WAIT_LOOP;\n"
		"DONE:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <typename V, int T>
__device__ __forceinline__ void res_gather(const V *s_val, const uint16_t *s_col, const double *s_x, const int *s_gid, const double4 *x,
	int n_own, int r0, int r1, int lane, double &sx, double &sy, double &sz)
{
	sx = 0; sy = 0; sz = 0;
#pragma unroll 4
	for (int r = r0; r < r1; ++r) {
		const int c = s_col[r * 32 + lane];
		const double a = (double)s_val[r * 32 + lane];
		double x0, x1, x2;
		if (c < n_own) { x0 = s_x[3 * c]; x1 = s_x[3 * c + 1]; x2 = s_x[3 * c + 2]; }
		else { double4 t = ld_node_cg(&x[s_gid[c]]); x0 = t.x; x1 = t.y; x2 = t.z; }
		sx += a * x0; sy += a * x1; sz += a * x2;
	}
#pragma unroll
	for (int o = 1; o < T; o <<= 1) {
		sx += __shfl_xor_sync(0xffffffffu, sx, o);
		sy += __shfl_xor_sync(0xffffffffu, sy, o);
		sz += __shfl_xor_sync(0xffffffffu, sz, o);
	}
}

// T lanes cooperate on one node (T = 1: one node per lane, no shuffles, every lane does an update;
// T = 4: shorter per-lane chains but only a quarter of the lanes own a node).
template <typename V, int T>
__global__ void __launch_bounds__(ADMMB200_RES_THREADS, 1) mcgs_resident_kernel(McgsResParams R)
{
	constexpr int G = 32 / T;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ double red[32];
	__shared__ __align__(8) uint64_t tma_bar;
	__shared__ int s_nbr[192];
	__shared__ int s_decision;
	const McgsParams &P = R.base;
	const PartDesc d = R.parts[blockIdx.x];
	const int tid = threadIdx.x, lane = tid & 31, sub = lane % T, grp = lane / T, warp = tid >> 5, n_warps = blockDim.x >> 5;

	// shared-memory layout: must match ResidentPlan::layout
	size_t o = 0;
	auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
	double *s_x = (double *)(smem + take(sizeof(double) * 3 * (size_t)d.n_own));
	V *s_val = (V *)(smem + take(sizeof(V) * 32 * (size_t)d.n_rows));
	uint16_t *s_col = (uint16_t *)(smem + take(sizeof(uint16_t) * 32 * (size_t)d.n_rows));
	int *s_gid = (int *)(smem + take(sizeof(int) * ((size_t)d.n_own + d.n_halo)));
	int *s_srow = (int *)(smem + take(sizeof(int) * ((size_t)d.n_slices + 1)));
	short *s_snode = (short *)(smem + take(sizeof(short) * (size_t)G * d.n_slices));
	int *s_cslice = (int *)(smem + take(sizeof(int) * (2 * (size_t)P.n_colors + 1)));

	// ---- stage the part: matrix by TMA bulk copy, the small index arrays and x by plain loads ----
	const uint32_t val_bytes = (uint32_t)(sizeof(V) * 32 * (size_t)d.n_rows), col_bytes = (uint32_t)(sizeof(uint16_t) * 32 * (size_t)d.n_rows);
	if (tid == 0) mbar_init(&tma_bar, 1);
	__syncthreads();
	if (tid == 0 && d.n_rows > 0) {
		mbar_expect_tx(&tma_bar, val_bytes + col_bytes);
		const unsigned char *gv = (const unsigned char *)R.val + (size_t)d.ent_off * sizeof(V);
		const unsigned char *gc = (const unsigned char *)(R.col + d.ent_off);
		const uint32_t chunk = 32768; // bytes per bulk copy
		for (uint32_t at = 0; at < val_bytes; at += chunk) bulk_g2s((unsigned char *)s_val + at, gv + at, min(chunk, val_bytes - at), &tma_bar);
		for (uint32_t at = 0; at < col_bytes; at += chunk) bulk_g2s((unsigned char *)s_col + at, gc + at, min(chunk, col_bytes - at), &tma_bar);
	}
	for (int i = tid; i < d.n_own + d.n_halo; i += blockDim.x) s_gid[i] = R.gid[d.gid_off + i];
	for (int i = tid; i <= d.n_slices; i += blockDim.x) s_srow[i] = R.slice_row[d.slice_off + i];
	for (int i = tid; i < G * d.n_slices; i += blockDim.x) s_snode[i] = R.slice_node[d.snode_off + i];
	for (int i = tid; i <= 2 * P.n_colors; i += blockDim.x) s_cslice[i] = R.color_slice[d.cslice_off + i];
	for (int i = tid; i < d.n_nbr; i += blockDim.x) s_nbr[i] = R.nbr[d.nbr_off + i];
	for (int l = tid; l < d.n_own; l += blockDim.x) {
		double4 xv = P.x[R.gid[d.gid_off + l]];
		s_x[3 * l] = xv.x; s_x[3 * l + 1] = xv.y; s_x[3 * l + 2] = xv.z;
	}
	if (d.n_rows > 0) mbar_wait(&tma_bar, 0);
	__syncthreads();

	unsigned int bar_target = 0;
	const bool check = P.tol2 > 0.0;
	const double omega = P.omega, one_m_omega = 1.0 - P.omega, lb_scale = 1.0 / P.omega - 1.0;
	if (check) {
		// b_norm = |b|^2 (src/NodalMultiColorGS.hpp:92)
		double acc = 0;
		for (int i = blockIdx.x * blockDim.x + tid; i < P.n_nodes; i += gridDim.x * blockDim.x) {
			double4 bi = ld_node(&P.b[i]);
			acc += bi.x * bi.x + bi.y * bi.y + bi.z * bi.z;
		}
		double s = block_sum(acc, red);
		if (tid == 0) atomicAdd(&P.resid[0], s);
		grid_barrier(P.barrier, bar_target, gridDim.x); // the only grid-wide barrier of a solve
	}
	const double thresh = check ? 4.0 * P.tol2 * __ldcg(&P.resid[0]) : 0.0;

	// One slice = G nodes of one colour: gather, SOR update, write back.  `to_global`: boundary nodes are
	// read by other parts and go to global memory at once; interior ones only at the end of the solve.
	double lb = 0;
	long long ps_meta = 0, ps_gather = 0, ps_tail = 0, ps_n = 0;
	auto do_slice = [&](int sl, bool to_global, bool last) {
		long long q0 = 0, q1 = 0, q2 = 0;
		if (R.prof) q0 = clock64();
		const int l = s_snode[sl * G + grp];
		const bool owner = (sub == 0 && l >= 0);
		double4 bi = make_double4(0, 0, 0, 0);
		double a0 = 1, a1 = 1, a2 = 1;
		int ps = -1, node = 0;
		if (owner) {
			node = s_gid[l];
			bi = ld_node(&P.b[node]);
			a0 = __ldg(&P.diag[3 * node]); a1 = __ldg(&P.diag[3 * node + 1]); a2 = __ldg(&P.diag[3 * node + 2]);
			if (P.has_pins) ps = __ldg(&P.pin_slot[node]);
		}
		double sx, sy, sz;
		if (R.prof) q1 = clock64();
		res_gather<V, T>(s_val, s_col, s_x, s_gid, P.x, d.n_own, s_srow[sl], s_srow[sl + 1], lane, sx, sy, sz);
		if (R.prof) q2 = clock64();
		if (owner) {
			double nx[3];
			if (ps >= 0) { nx[0] = P.pin_pos[3 * ps]; nx[1] = P.pin_pos[3 * ps + 1]; nx[2] = P.pin_pos[3 * ps + 2]; }
			else {
				const double xo[3] = {s_x[3 * l], s_x[3 * l + 1], s_x[3 * l + 2]};
				// segment_update (src/NodalMultiColorGS.hpp:180-215)
				double gs[3] = {(bi.x - sx) / a0, (bi.y - sy) / a1, (bi.z - sz) / a2};
				nx[0] = one_m_omega * xo[0] + omega * gs[0]; nx[1] = one_m_omega * xo[1] + omega * gs[1]; nx[2] = one_m_omega * xo[2] + omega * gs[2];
				bool hit = false;
				if (P.n_obstacles > 0) hit = mcgs_collide(P.obs, P.n_obstacles, gs, nx);
				if (last && !hit) {
					double rx = a0 * lb_scale * (nx[0] - xo[0]), ry = a1 * lb_scale * (nx[1] - xo[1]), rz = a2 * lb_scale * (nx[2] - xo[2]);
					lb += rx * rx + ry * ry + rz * rz;
				}
			}
			s_x[3 * l] = nx[0]; s_x[3 * l + 1] = nx[1]; s_x[3 * l + 2] = nx[2];
			if (to_global) st_node(&P.x[node], nx[0], nx[1], nx[2]);
		}
		if (R.prof) { long long q3 = clock64(); ps_meta += q1 - q0; ps_gather += q2 - q1; ps_tail += q3 - q2; ps_n += 1; }
	};

	// Warp roles: warps [0, n_bwarps) own the boundary slices -- the chain "wait for the neighbours ->
	// update -> publish" that limits the pass rate -- the others update the interior concurrently.  Both
	// meet at one __syncthreads per pass: pass c+1 reads what both groups wrote in pass c.
	const int n_bwarps = n_warps / 2, n_iwarps = n_warps - n_bwarps;
	const bool bwarp = warp < n_bwarps;
	int it = 0;
	unsigned int epoch = 0;
	long long pw = 0, pc = 0, pp = 0, pi = 0;
	const long long t_begin = R.prof ? clock64() : 0;
	for (; it < P.iters; ++it) {
		lb = 0;
		for (int color = 0; color < P.n_colors; ++color) {
			const int s0 = s_cslice[2 * color], s1 = s_cslice[2 * color + 1], s2 = s_cslice[2 * color + 2];
			const bool last = check && (color == P.n_colors - 1);
			long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
			if (R.prof) t0 = clock64();
			if (bwarp) {
				if (epoch > 0) part_wait(R.part_epoch, s_nbr, d.n_nbr, epoch, 32 * n_bwarps);
				if (R.prof) t1 = clock64();
				for (int sl = s1 + warp; sl < s2; sl += n_bwarps) do_slice(sl, true, last);
				if (R.prof) t2 = clock64();
			} else {
				// interior nodes: nobody else reads or writes them and they read only this part's values
				for (int sl = s0 + (warp - n_bwarps); sl < s1; sl += n_iwarps) do_slice(sl, false, last);
				if (R.prof) t1 = clock64();
			}
			__syncthreads();
			if (R.prof) t3 = clock64();
			part_publish(R.part_epoch, ++epoch);
			if (R.prof && tid == 0) { long long t4 = clock64(); pw += t1 - t0; pc += t2 - t1; pp += t4 - t3; }
			if (R.prof && tid == 32 * n_bwarps) pi += t1 - t0;
		}
		if (check) {
			// Decide "converged?" without a grid barrier in the common case: any part whose own rows
			// already prove |b - A x|^2 >= 4 tol^2 |b|^2 raises sweep_flag; a part continues as soon as it
			// sees the flag.  Only if nobody can prove it do all parts meet (sweep_arrive == grid) and look
			// at the summed bound, then at the exact residual.
			double s = block_sum(lb, red);
			if (tid == 0) {
				if (s >= thresh) R.sweep_flag[it] = 1u;
				else if (s > 0.0) atomicAdd(&P.resid_lb[it], s);
				__threadfence();
				atomicAdd(&R.sweep_arrive[it], 1u);
				int decision = -1;
				while (decision < 0) {
					if (ld_relaxed_u32(&R.sweep_flag[it]) != 0u) decision = 1;
					else if (ld_relaxed_u32(&R.sweep_arrive[it]) == gridDim.x) {
						fence_acq_rel_gpu();
						decision = (ld_relaxed_u32(&R.sweep_flag[it]) != 0u || __ldcg(&P.resid_lb[it]) >= thresh) ? 1 : 0;
					}
				}
				s_decision = decision;
			}
			__syncthreads();
			const bool proven_unconverged = s_decision == 1;
			__syncthreads();
			if (!proven_unconverged) {
				const double b2 = __ldcg(&P.resid[0]);
				// every part is here (sweep_arrive == grid) and all boundary values of this sweep are published
				grid_barrier(P.barrier, bar_target, gridDim.x);
				// exact residual b - A x (src/NodalMultiColorGS.hpp:136-139)
				double acc = 0;
				for (int sl = warp; sl < d.n_slices; sl += n_warps) {
					const int l = s_snode[sl * G + grp];
					double sx, sy, sz;
					res_gather<V, T>(s_val, s_col, s_x, s_gid, P.x, d.n_own, s_srow[sl], s_srow[sl + 1], lane, sx, sy, sz);
					if (sub == 0 && l >= 0) {
						const int node = s_gid[l];
						double4 bi = ld_node(&P.b[node]);
						double rx = bi.x - (sx + P.diag[3 * node] * s_x[3 * l]);
						double ry = bi.y - (sy + P.diag[3 * node + 1] * s_x[3 * l + 1]);
						double rz = bi.z - (sz + P.diag[3 * node + 2] * s_x[3 * l + 2]);
						acc += rx * rx + ry * ry + rz * rz;
					}
				}
				double sres = block_sum(acc, red);
				if (tid == 0) atomicAdd(&P.resid[1 + it], sres);
				grid_barrier(P.barrier, bar_target, gridDim.x);
				double r2 = __ldcg(&P.resid[1 + it]);
				if (r2 / b2 < P.tol2) break;
			}
		}
	}
	// the interior nodes have only been updated in shared memory: hand the whole part back
	__syncthreads();
	for (int l = tid; l < d.n_own; l += blockDim.x) st_node(&P.x[s_gid[l]], s_x[3 * l], s_x[3 * l + 1], s_x[3 * l + 2]);
	if (blockIdx.x == 0 && tid == 0) *P.iters_done = it;
	if (R.prof && tid == 0) {
		unsigned long long *q = R.prof + 16 * blockIdx.x;
		q[0] = (unsigned long long)pw; q[1] = (unsigned long long)pc; q[2] = (unsigned long long)pp; q[3] = (unsigned long long)(clock64() - t_begin);
		q[5] = (unsigned long long)ps_meta; q[6] = (unsigned long long)ps_gather; q[7] = (unsigned long long)ps_tail; q[8] = (unsigned long long)ps_n;
	}
	if (R.prof && tid == 32 * n_bwarps) {
		unsigned long long *q = R.prof + 16 * blockIdx.x;
		q[4] = (unsigned long long)pi;
		q[9] = (unsigned long long)ps_meta; q[10] = (unsigned long long)ps_gather; q[11] = (unsigned long long)ps_tail; q[12] = (unsigned long long)ps_n;
	}
}

} // namespace admmb200
