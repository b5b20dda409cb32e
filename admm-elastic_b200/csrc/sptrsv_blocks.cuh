// sptrsv_blocks.cuh -- prefactored L D L^T solve on the GPU by BLOCKS (supernodes), 3 right-hand sides at once.
//
// Replaces LDLTSolver::solve = Eigen SimplicialLDLT::solve (src/LinearSolver.hpp:87-90): x = P^T L^-T D^-1 L^-1 P b, and is
// the inner solve of UzawaCG (src/UzawaCG.hpp:83-118).  Plan: ldlt_blocks.hpp -- the columns of L are cut into blocks whose
// diagonal part is inverted on the host, so a level of the block tree costs
//     gather:  t_i = b_i - sum_{j left of i's block} L_ij y_j        one item per row, lanes stride over the row (coalesced)
//     dense:   y_i = t_i + sum_{k < i in the block} Linv_ik t_k      one item per row, the block's t is contiguous
// with a grid barrier after each: ~2 x 15 barriers per direction instead of one per dependency level (2 415 on the cloth).
// One persistent cooperative launch per solve; A = L_scalar (x) I3, so x, y, z go through together as double4.
#pragma once
#include "sptrsv.cuh"

namespace admmb200 {

struct LdltBlkParams {
	int n, n_levels, cut;            // block levels; levels below `cut` belong to the bottom forest (CTA-local, no grid barrier)
	const int *perm;                 // [n] perm[new] = old
	const int *blk_of, *blk_c0;      // [n], [n_blocks + 1]
	const long long *inv_off;        // [n_blocks]
	const double *inv, *invT;        // packed strictly-lower inverse rows / transposed rows
	const int *lev_ptr, *rows;       // rows (= columns) ordered by the level of their block
	const int *lanes;                // [4 * n_levels] threads per row: forward gather, forward dense, backward gather, backward dense
	const int *f_rowptr, *f_cols;    // entries of a row LEFT of its block (CSR)
	const double *f_vals;
	const int *b_colptr, *b_rows;    // entries of a column BELOW its block (CSC)
	const double *b_vals;
	const int *seg_ptr, *seg_begin, *seg_end, *seg_level; // the forest: per CTA its segments (ranges of `rows`) in forward order
	const int4 *desc_fg, *desc_bg, *desc_d;   // per position k of `rows`: see LdltBlockPlan::build_descriptors
	const long long *desc_off;
	const double *D;
	double4 *t, *y;                  // work, permuted numbering
	const double4 *b;                // node order
	double4 *x;                      // node order, out
	unsigned int *barrier;
	const int *active;               // NULL, or a device flag: 0 = skip this solve (uzawa.cuh: the CG loop has already ended)
	unsigned long long *prof;        // NULL, or [4 * n_levels + 8] nanosecond stamps of CTA 0 (ADMM_B200_LDLT_PROF=1, tools/ldlt_prof.py)
};
__device__ __forceinline__ void blk_stamp(const LdltBlkParams &P, int slot) {
	if (P.prof && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); P.prof[slot] = t; }
}

// Sum over the T threads of a group (T a power of two, 1..1024; groups are aligned, so one never straddles a CTA).
// T <= 32: shuffles.  T > 32: every warp reduces, lane 0 leaves its partial sum in shared memory, the group's first warp
// adds them up (two CTA barriers -- every thread of the CTA calls this the same number of times).  Valid in sub == 0.
__device__ __forceinline__ void blk_reduce(double &sx, double &sy, double &sz, int T, double *s_part)
{
	const int w = T < 32 ? T : 32;
	for (int o = w >> 1; o > 0; o >>= 1) {
		sx += __shfl_xor_sync(0xffffffffu, sx, o);
		sy += __shfl_xor_sync(0xffffffffu, sy, o);
		sz += __shfl_xor_sync(0xffffffffu, sz, o);
	}
	if (T <= 32) return;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpg = T >> 5; // warps per group
	__syncthreads();
	if (lane == 0) { s_part[3 * warp] = sx; s_part[3 * warp + 1] = sy; s_part[3 * warp + 2] = sz; }
	__syncthreads();
	if ((warp & (wpg - 1)) == 0 && lane == 0) {
		for (int k = 1; k < wpg; ++k) { sx += s_part[3 * (warp + k)]; sy += s_part[3 * (warp + k) + 1]; sz += s_part[3 * (warp + k) + 2]; }
	}
}

// dot product of a sparse / dense row with double4 vectors: 4 independent loads in flight per lane (the rows of the top
// separators have thousands of entries and every gather is an L2 round trip)
template <bool INDEXED>
__device__ __forceinline__ void blk_row_dot(const double *__restrict__ vals, const int *__restrict__ idx, const double4 *vec, int q0, int q1, int sub, int T,
	double &sx, double &sy, double &sz)
{
	int q = q0 + sub;
	for (; q + 3 * T < q1; q += 4 * T) {
		double a[4]; int c[4]; double4 v[4];
#pragma unroll
		for (int u = 0; u < 4; ++u) { a[u] = __ldg(&vals[q + u * T]); c[u] = INDEXED ? __ldg(&idx[q + u * T]) : q + u * T; }
#pragma unroll
		for (int u = 0; u < 4; ++u) v[u] = ld_node_cg(&vec[c[u]]);
#pragma unroll
		for (int u = 0; u < 4; ++u) { sx += a[u] * v[u].x; sy += a[u] * v[u].y; sz += a[u] * v[u].z; }
	}
	for (; q < q1; q += T) {
		const double a = __ldg(&vals[q]);
		const double4 v = ld_node_cg(&vec[INDEXED ? __ldg(&idx[q]) : q]);
		sx += a * v.x; sy += a * v.y; sz += a * v.z;
	}
}

// One phase over the rows rows[k0, k1), shared by `n_thr` threads of which this one is number `t`:
//   PHASE 0  forward gather   t_i = b_i - sum_{j left of the block} L_ij y_j
//   PHASE 1  forward dense    y_i = t_i + sum_{k < i in the block} Linv_ik t_k
//   PHASE 2  backward gather  t_j = y_j / D_j - sum_{i below the block} L_ij x_i
//   PHASE 3  backward dense   x_j = t_j + sum_{k > j in the block} Linv_kj t_k
// Every thread of the CTA runs the same trip count (blk_reduce may hold CTA barriers).
// The descriptor of this thread's FIRST row of a phase.  It does not depend on anything the solve computes, so the caller
// asks for it before the barrier that precedes the phase: one L2 round trip less on the chain after the barrier.
template <int PHASE>
__device__ __forceinline__ int4 blk_prefetch(const LdltBlkParams &P, int k0, int k1, int T, int t)
{
	const int k = k0 + t / T;
	if (k >= k1) return make_int4(0, 0, 0, 0);
	return __ldg(PHASE == 0 ? &P.desc_fg[k] : (PHASE == 2 ? &P.desc_bg[k] : &P.desc_d[k]));
}

template <int PHASE>
__device__ __forceinline__ void blk_phase(const LdltBlkParams &P, const double4 *b, double4 *x, int k0, int k1, int T, int t, int n_thr, double *s_part, int4 pref)
{
	const int sub = t & (T - 1), group = t / T, n_groups = n_thr / T;
	for (int kb = k0; kb < k1; kb += n_groups) {
		const int k = kb + group;
		const bool act = k < k1;
		double sx = 0, sy = 0, sz = 0;
		int4 ds = make_int4(0, 0, 0, 0);
		double dj = 1.0;
		double4 own = make_double4(0, 0, 0, 0); // the row's own right-hand side / intermediate value: asked for before the dot product
		if (act) {
			// one 16-byte descriptor instead of the chain rows[k] -> rowptr / block tables -> entries
			ds = kb == k0 ? pref : __ldg(PHASE == 0 ? &P.desc_fg[k] : (PHASE == 2 ? &P.desc_bg[k] : &P.desc_d[k]));
			if (sub == 0) own = PHASE == 0 ? ld_node_cg(&b[ds.w]) : (PHASE == 2 ? ld_node_cg(&P.y[ds.x]) : ld_node_cg(&P.t[ds.w]));
			if (PHASE == 0) blk_row_dot<true>(P.f_vals, P.f_cols, P.y, ds.y, ds.z, sub, T, sx, sy, sz);
			else if (PHASE == 2) { if (sub == 0) dj = __ldg(&P.D[ds.x]); blk_row_dot<true>(P.b_vals, P.b_rows, P.y, ds.y, ds.z, sub, T, sx, sy, sz); } // rows below the block: already final
			else {
				const long long off = __ldg(&P.desc_off[k]);
				const int c0 = ds.x, r = ds.y, s = ds.z;
				if (PHASE == 1) blk_row_dot<false>(P.inv + off + (long long)r * (r - 1) / 2, nullptr, P.t + c0, 0, r, sub, T, sx, sy, sz);
				else blk_row_dot<false>(P.invT + off + (long long)r * (s - 1) - (long long)r * (r - 1) / 2, nullptr, P.t + c0 + r + 1, 0, s - r - 1, sub, T, sx, sy, sz);
			}
		}
		blk_reduce(sx, sy, sz, T, s_part);
		if (act && sub == 0) {
			if (PHASE == 0) st_node(&P.t[ds.x], own.x - sx, own.y - sy, own.z - sz);
			else if (PHASE == 1) st_node(&P.y[ds.w], own.x + sx, own.y + sy, own.z + sz);
			else if (PHASE == 2) st_node(&P.t[ds.x], own.x / dj - sx, own.y / dj - sy, own.z / dj - sz);
			else {
				const double rx = own.x + sx, ry = own.y + sy, rz = own.z + sz;
				st_node(&P.y[ds.w], rx, ry, rz);
				st_node(&x[__ldg(&P.perm[ds.w])], rx, ry, rz);
			}
		}
	}
}

// The whole solve x = P^T L^-T D^-1 L^-1 P b as a device function: the standalone kernel below and the persistent UzawaCG
// kernel (uzawa_blocks.cuh) call it.  Every thread of the (cooperative) grid must call it; it ends with a grid barrier.
// The bottom forest (levels < cut) is walked CTA by CTA with CTA barriers only -- its subtrees are independent of each
// other -- so only the 2 x (n_levels - cut) phases above the cut cost a grid barrier per direction.
__device__ __forceinline__ void ldlt_blocks_solve(const LdltBlkParams &P, const double4 *b, double4 *x, unsigned int &bar_target, double *s_part)
{
	const int ltid = threadIdx.x, lthr = blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x, gthr = gridDim.x * blockDim.x;
	const int sg0 = __ldg(&P.seg_ptr[blockIdx.x]), sg1 = __ldg(&P.seg_ptr[blockIdx.x + 1]);
	const int4 none = make_int4(0, 0, 0, 0);
	// a phase's first descriptor is always asked for BEFORE the barrier in front of the phase (blk_prefetch)
	struct Seg { int k0, k1, lv; };
	auto seg = [&](int sg) { Seg g; g.k0 = __ldg(&P.seg_begin[sg]); g.k1 = __ldg(&P.seg_end[sg]); g.lv = __ldg(&P.seg_level[sg]); return g; };
	auto lev = [&](int lv) { Seg g; g.k0 = __ldg(&P.lev_ptr[lv]); g.k1 = __ldg(&P.lev_ptr[lv + 1]); g.lv = lv; return g; };
	const bool top = P.cut < P.n_levels;
	blk_stamp(P, 0);
	// ---------------- forward: y = L^-1 P b ----------------
	{
		Seg g = sg0 < sg1 ? seg(sg0) : Seg{0, 0, 0};
		int4 pf = sg0 < sg1 ? blk_prefetch<0>(P, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv]), ltid) : none;
		for (int sg = sg0; sg < sg1; ++sg) {
			const int T1 = __ldg(&P.lanes[4 * g.lv + 1]);
			blk_phase<0>(P, b, x, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv]), ltid, lthr, s_part, pf);
			pf = blk_prefetch<1>(P, g.k0, g.k1, T1, ltid);
			__syncthreads();
			blk_phase<1>(P, b, x, g.k0, g.k1, T1, ltid, lthr, s_part, pf);
			if (sg + 1 < sg1) { g = seg(sg + 1); pf = blk_prefetch<0>(P, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv]), ltid); }
			__syncthreads();
		}
	}
	blk_stamp(P, 1); // own forest done
	{
		Seg g = top ? lev(P.cut) : Seg{0, 0, 0};
		int4 pf = top ? blk_prefetch<0>(P, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv]), gtid) : none;
		if (P.cut > 0) grid_barrier(P.barrier, bar_target, gridDim.x);
		blk_stamp(P, 2); // everybody's forest done
		for (int lv = P.cut; lv < P.n_levels; ++lv) {
			const int T1 = __ldg(&P.lanes[4 * lv + 1]);
			blk_phase<0>(P, b, x, g.k0, g.k1, __ldg(&P.lanes[4 * lv]), gtid, gthr, s_part, pf);
			pf = blk_prefetch<1>(P, g.k0, g.k1, T1, gtid);
			grid_barrier(P.barrier, bar_target, gridDim.x);
			blk_stamp(P, 8 + 4 * lv);
			blk_phase<1>(P, b, x, g.k0, g.k1, T1, gtid, gthr, s_part, pf);
			if (lv + 1 < P.n_levels) { g = lev(lv + 1); pf = blk_prefetch<0>(P, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv]), gtid); }
			else pf = blk_prefetch<2>(P, g.k0, g.k1, __ldg(&P.lanes[4 * lv + 2]), gtid); // the backward sweep starts on this level
			grid_barrier(P.barrier, bar_target, gridDim.x);
			blk_stamp(P, 8 + 4 * lv + 1);
		}
		// ---------------- backward: x = P^T L^-T D^-1 y, the same levels in reverse (an ancestor sits on a higher level) ----------------
		for (int lv = P.n_levels - 1; lv >= P.cut; --lv) {
			const int T3 = __ldg(&P.lanes[4 * lv + 3]);
			blk_phase<2>(P, b, x, g.k0, g.k1, __ldg(&P.lanes[4 * lv + 2]), gtid, gthr, s_part, pf);
			pf = blk_prefetch<3>(P, g.k0, g.k1, T3, gtid);
			grid_barrier(P.barrier, bar_target, gridDim.x);
			blk_stamp(P, 8 + 4 * lv + 2);
			blk_phase<3>(P, b, x, g.k0, g.k1, T3, gtid, gthr, s_part, pf);
			if (lv > P.cut) { g = lev(lv - 1); pf = blk_prefetch<2>(P, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv + 2]), gtid); }
			else if (sg0 < sg1) { g = seg(sg1 - 1); pf = blk_prefetch<2>(P, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv + 2]), ltid); } // own forest next
			grid_barrier(P.barrier, bar_target, gridDim.x);
			blk_stamp(P, 8 + 4 * lv + 3);
		}
		blk_stamp(P, 3); // top done
		if (!top && sg0 < sg1) { g = seg(sg1 - 1); pf = blk_prefetch<2>(P, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv + 2]), ltid); }
		for (int sg = sg1 - 1; sg >= sg0; --sg) {
			const int T3 = __ldg(&P.lanes[4 * g.lv + 3]);
			blk_phase<2>(P, b, x, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv + 2]), ltid, lthr, s_part, pf);
			pf = blk_prefetch<3>(P, g.k0, g.k1, T3, ltid);
			__syncthreads();
			blk_phase<3>(P, b, x, g.k0, g.k1, T3, ltid, lthr, s_part, pf);
			if (sg > sg0) { g = seg(sg - 1); pf = blk_prefetch<2>(P, g.k0, g.k1, __ldg(&P.lanes[4 * g.lv + 2]), ltid); }
			__syncthreads();
		}
	}
	blk_stamp(P, 4); // own forest done (backward)
	if (P.cut > 0) grid_barrier(P.barrier, bar_target, gridDim.x); // callers read x right away
	blk_stamp(P, 5);
}

__global__ void __launch_bounds__(1024, 1) ldlt_blocks_kernel(LdltBlkParams P)
{
	__shared__ double s_part[3 * 32];
	unsigned int bar_target = 0;
	if (P.active && *P.active == 0) return; // the same for every block: no barrier is left waiting
	ldlt_blocks_solve(P, P.b, P.x, bar_target, s_part);
}

} // namespace admmb200
