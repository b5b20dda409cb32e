// sptrsv.cuh -- prefactored sparse L D L^T solve on the GPU, 3 right-hand sides at once.
//
// Replaces LDLTSolver::solve = Eigen SimplicialLDLT::solve (src/LinearSolver.hpp:87-90;
// deps/Eigen3/Eigen/src/SparseCholesky/SimplicialCholesky.h:156-180): x = P^T L^-T D^-1 L^-1 P b.
// Because A = L_scalar (x) I3 the n x n scalar factor is applied to the x, y and z
// components together (one double4 per node).
//
// Level-scheduled gather form, one persistent cooperative launch per solve:
//   forward : y_i = b_p(i) - sum_{j<i} L_ij y_j       rows grouped by dependency level
//   diagonal: y_i /= D_i                               fused into the last touch of row i
//   backward: x_i = y_i - sum_{j>i} L_ji x_j           levels of the transposed DAG
// Rows of one level are independent; a grid barrier separates levels.
#pragma once
#include "kernels.cuh"

namespace admmb200 {

struct LdltParams {
	int n;
	int n_levels_fwd, n_levels_bwd;
	const int *perm;          // [n] perm[new] = old
	// forward: rows ordered by level
	const int *fwd_level_ptr; // [n_levels_fwd+1] into fwd_rows
	const int *fwd_rows;      // [n] row ids (permuted numbering)
	const int *fwd_rowptr;    // [n+1] CSR of strictly-lower L
	const int *fwd_cols;
	const double *fwd_vals;
	const int *bwd_level_ptr;
	const int *bwd_rows;
	const int *bwd_rowptr;    // [n+1] CSR of L^T (= CSC of L), strictly upper
	const int *bwd_cols;
	const double *bwd_vals;
	const double *dinv_unused;
	const double *D;          // [n]
	double4 *y;               // [n] work (permuted numbering)
	const double4 *b;         // [n] node order
	double4 *x;               // [n] node order, out
	unsigned int *barrier;
	const int *active;        // NULL, or a device flag: 0 = skip this solve (uzawa.cuh: the CG loop has already ended)
};

// T lanes cooperate on one row
template <int T>
__global__ void __launch_bounds__(512, 1) ldlt_solve_kernel(LdltParams P)
{
	const int lane = threadIdx.x & 31;
	const int sub = lane % T;
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int group = tid / T;
	const int n_groups = (gridDim.x * blockDim.x) / T;
	unsigned int bar_target = 0;
	if (P.active && *P.active == 0) return; // the same for every block: no barrier is left waiting

	for (int lv = 0; lv < P.n_levels_fwd; ++lv) {
		const int k0 = P.fwd_level_ptr[lv], k1 = P.fwd_level_ptr[lv + 1];
		for (int kb = k0; kb < k1; kb += n_groups) {
			int k = kb + group;
			bool act = k < k1;
			int i = act ? P.fwd_rows[k] : 0;
			double sx = 0, sy = 0, sz = 0;
			if (act) {
				for (int q = P.fwd_rowptr[i] + sub; q < P.fwd_rowptr[i + 1]; q += T) {
					double a = __ldg(&P.fwd_vals[q]);
					double4 yj = ld_node_cg(&P.y[__ldg(&P.fwd_cols[q])]);
					sx += a * yj.x; sy += a * yj.y; sz += a * yj.z;
				}
			}
#pragma unroll
			for (int o = 1; o < T; o <<= 1) {
				sx += __shfl_xor_sync(0xffffffffu, sx, o);
				sy += __shfl_xor_sync(0xffffffffu, sy, o);
				sz += __shfl_xor_sync(0xffffffffu, sz, o);
			}
			if (act && sub == 0) {
				double4 bi = P.b[P.perm[i]];
				st_node(&P.y[i], bi.x - sx, bi.y - sy, bi.z - sz);
			}
		}
		grid_barrier(P.barrier, bar_target, gridDim.x);
	}
	for (int lv = 0; lv < P.n_levels_bwd; ++lv) {
		const int k0 = P.bwd_level_ptr[lv], k1 = P.bwd_level_ptr[lv + 1];
		for (int kb = k0; kb < k1; kb += n_groups) {
			int k = kb + group;
			bool act = k < k1;
			int i = act ? P.bwd_rows[k] : 0;
			double sx = 0, sy = 0, sz = 0;
			if (act) {
				for (int q = P.bwd_rowptr[i] + sub; q < P.bwd_rowptr[i + 1]; q += T) {
					double a = __ldg(&P.bwd_vals[q]);
					double4 xj = ld_node_cg(&P.y[__ldg(&P.bwd_cols[q])]);
					sx += a * xj.x; sy += a * xj.y; sz += a * xj.z;
				}
			}
#pragma unroll
			for (int o = 1; o < T; o <<= 1) {
				sx += __shfl_xor_sync(0xffffffffu, sx, o);
				sy += __shfl_xor_sync(0xffffffffu, sy, o);
				sz += __shfl_xor_sync(0xffffffffu, sz, o);
			}
			if (act && sub == 0) {
				// y_i currently holds the forward result; D^-1 is applied here, on its last use
				double4 yi = ld_node_cg(&P.y[i]);
				double d = P.D[i];
				double rx = yi.x / d - sx, ry = yi.y / d - sy, rz = yi.z / d - sz;
				st_node(&P.y[i], rx, ry, rz);
				st_node(&P.x[P.perm[i]], rx, ry, rz);
			}
		}
		grid_barrier(P.barrier, bar_target, gridDim.x);
	}
}

} // namespace admmb200
