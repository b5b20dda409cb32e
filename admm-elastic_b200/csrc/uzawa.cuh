// uzawa.cuh -- UzawaCG::solve (src/UzawaCG.hpp:57-125) with passive collisions, on the device.
//
// The reference builds the constraint matrix C from the passive hits of this ADMM iteration
// (Collider::detect with_passive, src/Collider.hpp:152-212; ConstraintSet::make_matrix,
// src/ConstraintSet.hpp:59-116: one row n^T per hit vertex, c = n . p, constraint_w = 1 for linsolver 2,
// src/Solver.cpp:239) and runs conjugate gradients on the Schur complement C A^-1 C^T, every product with
// A^-1 being one prefactored LDL^T solve.  Here C is never formed: a row is (vertex, normal, c), so
// C^T d is a scatter of 3-vectors and C q a gather; A^-1 is ldlt_solve_kernel (sptrsv.cuh) with 3
// right-hand sides.  The iteration count is data dependent (two `break`s), and no host round trip
// decides it: a device flag `active` turns the remaining launches of the fixed-length sequence into
// no-ops, and the count the reference returns is kept on the device.
//
//   ctl[0] rows   ctl[1] rows of the previous solve (the multipliers y are warm-started when the count
//   is unchanged, :69-74)   ctl[2] active   ctl[3] apply this iteration's x update   ctl[4] iter
#pragma once
#include "kernels.cuh"
#include <cfloat>

namespace admmb200 {

struct UzParams {
	int n;                 // nodes
	int n_obstacles;
	const Obstacle *obs;
	int *hv;               // [n] hit vertex per row
	double *hn, *hc;       // [3n] normals, [n] c = n . p
	double *y, *r, *d, *q3; // [n] multipliers, residual, direction, C q2
	int *ctl;              // [8]
	double *scal;          // [2] alpha
	double tol2;
};

// Collider::detect for passive objects: EVERY object lowers the payload (no early exit, unlike
// Collider::detect_passive inside the Gauss-Seidel sweep); hit if the smallest signed distance is < 0.
__device__ __forceinline__ bool uz_detect_node(const Obstacle *obs, int n_obstacles, const double *x, double *nrm, double *pt)
{
	double dx = 1.7976931348623157e308;
	for (int j = 0; j < n_obstacles; ++j) {
		const Obstacle o = obs[j];
		if (o.kind == 0) { // Floor (src/PassiveObject.hpp:32-45)
			const double d = x[1] - o.p[0];
			if (!(d > dx)) { dx = d; pt[0] = x[0]; pt[1] = o.p[0]; pt[2] = x[2]; nrm[0] = 0; nrm[1] = 1; nrm[2] = 0; }
		} else {           // Sphere (:47-64)
			double r0 = x[0] - o.p[0], r1 = x[1] - o.p[1], r2 = x[2] - o.p[2];
			const double len = sqrt(r0 * r0 + r1 * r1 + r2 * r2), d = len - o.p[3];
			if (!(d > dx)) {
				dx = d; r0 /= len; r1 /= len; r2 /= len;
				pt[0] = o.p[0] + r0 * o.p[3]; pt[1] = o.p[1] + r1 * o.p[3]; pt[2] = o.p[2] + r2 * o.p[3];
				nrm[0] = r0; nrm[1] = r1; nrm[2] = r2;
			}
		}
	}
	return dx < 0;
}

// One block: hits compacted in node order (the order the reference produces with one thread; it only
// matters for the warm start of y when the hit count happens to stay the same).
__global__ void __launch_bounds__(1024) uz_detect_kernel(UzParams U, const double4 *__restrict__ cx)
{
	__shared__ int s_cnt[1024];
	__shared__ int s_total, s_reset;
	const int tid = threadIdx.x, nt = blockDim.x;
	const int chunk = (U.n + nt - 1) / nt, i0 = tid * chunk, i1 = min(U.n, i0 + chunk);
	int cnt = 0;
	for (int i = i0; i < i1; ++i) {
		const double4 p = cx[i];
		const double x[3] = {p.x, p.y, p.z};
		double nrm[3], pt[3];
		if (uz_detect_node(U.obs, U.n_obstacles, x, nrm, pt)) ++cnt;
	}
	s_cnt[tid] = cnt;
	__syncthreads();
	// exclusive scan (Hillis-Steele on 1024 counters)
	for (int o = 1; o < nt; o <<= 1) {
		int v = tid >= o ? s_cnt[tid - o] : 0;
		__syncthreads();
		s_cnt[tid] += v;
		__syncthreads();
	}
	int at = s_cnt[tid] - cnt;
	if (tid == nt - 1) {
		const int rows = s_cnt[tid];
		s_total = rows;
		s_reset = (U.ctl[1] != rows); // y = 0 unless the row count is unchanged (src/UzawaCG.hpp:74)
		U.ctl[0] = rows; U.ctl[1] = rows; U.ctl[2] = rows > 0; U.ctl[3] = 0; U.ctl[4] = 0;
	}
	for (int i = i0; i < i1; ++i) {
		const double4 p = cx[i];
		const double x[3] = {p.x, p.y, p.z};
		double nrm[3], pt[3];
		if (uz_detect_node(U.obs, U.n_obstacles, x, nrm, pt)) {
			U.hv[at] = i;
			U.hn[3 * at] = nrm[0]; U.hn[3 * at + 1] = nrm[1]; U.hn[3 * at + 2] = nrm[2];
			U.hc[at] = nrm[0] * pt[0] + nrm[1] * pt[1] + nrm[2] * pt[2];
			++at;
		}
	}
	__syncthreads();
	if (s_reset) for (int k = tid; k < s_total; k += nt) U.y[k] = 0.0;
}

// out = in (+ sign * C^T w when w != NULL): the node-sized part
__global__ void uz_copy_kernel(int n, const double4 *__restrict__ in, double4 *__restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { const double4 v = in ? in[i] : make_double4(0, 0, 0, 0); st_node(&out[i], v.x, v.y, v.z); }
}
// q[hv[k]] += sign * n_k w_k  (a vertex is hit at most once, so rows never collide)
__global__ void uz_scatter_kernel(UzParams U, const double *__restrict__ w, double sign, double4 *__restrict__ q, int need_active)
{
	if (need_active && U.ctl[2] == 0) return;
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= U.ctl[0]) return;
	const int v = U.hv[k];
	const double a = sign * w[k];
	double4 t = q[v];
	st_node(&q[v], t.x + a * U.hn[3 * k], t.y + a * U.hn[3 * k + 1], t.z + a * U.hn[3 * k + 2]);
}
// r = C x - c, d = r   (src/UzawaCG.hpp:86-87)
__global__ void uz_init_kernel(UzParams U, const double4 *__restrict__ x)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= U.ctl[0]) return;
	const double4 p = x[U.hv[k]];
	const double r = U.hn[3 * k] * p.x + U.hn[3 * k + 1] * p.y + U.hn[3 * k + 2] * p.z - U.hc[k];
	U.r[k] = r; U.d[k] = r;
}

__device__ __forceinline__ double uz_block_sum(double v, double *red)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0;
	for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
	return s; // every thread
}

// The row-sized part of one CG iteration (:93-118), one block.  q2 = A^-1 C^T d has just been solved.
__global__ void __launch_bounds__(1024) uz_step_kernel(UzParams U, const double4 *__restrict__ q2)
{
	__shared__ double red[32];
	const int tid = threadIdx.x, nt = blockDim.x, rows = U.ctl[0];
	if (U.ctl[2] == 0) { if (tid == 0) U.ctl[3] = 0; return; }
	double denom = 0, dr = 0;
	for (int k = tid; k < rows; k += nt) {
		const double4 p = q2[U.hv[k]];
		const double q3 = U.hn[3 * k] * p.x + U.hn[3 * k + 1] * p.y + U.hn[3 * k + 2] * p.z;
		U.q3[k] = q3;
		denom += U.d[k] * q3; dr += U.d[k] * U.r[k];
	}
	denom = uz_block_sum(denom, red);
	dr = uz_block_sum(dr, red);
	if (fabs(denom) < DBL_MIN) { if (tid == 0) { U.ctl[2] = 0; U.ctl[3] = 0; } return; } // is_zero(denom): break
	const double alpha = dr / denom;
	double rr = 0, rq = 0;
	for (int k = tid; k < rows; k += nt) {
		U.y[k] += alpha * U.d[k];
		const double r = U.r[k] - alpha * U.q3[k];
		U.r[k] = r;
		rr += r * r; rq += r * U.q3[k];
	}
	rr = uz_block_sum(rr, red);
	rq = uz_block_sum(rq, red);
	if (tid == 0) { U.scal[0] = alpha; U.ctl[3] = 1; } // x -= alpha q2 happens before either of the next two exits
	if (rr < U.tol2) { if (tid == 0) U.ctl[2] = 0; return; }
	// (the second is_zero(denom) test of the reference looks at the same number)
	const double beta = rq / denom;
	for (int k = tid; k < rows; k += nt) U.d[k] = U.r[k] - beta * U.d[k];
	if (tid == 0) U.ctl[4] += 1;
}

// x -= alpha q2
__global__ void uz_axpy_kernel(int n, const int *__restrict__ ctl, const double *__restrict__ scal, const double4 *__restrict__ q2, double4 *__restrict__ x)
{
	if (ctl[3] == 0) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const double a = scal[0];
	const double4 p = x[i], q = q2[i];
	st_node(&x[i], p.x - a * q.x, p.y - a * q.y, p.z - a * q.z);
}

// return value of solve(): 1 when C is empty (:78-81), else the number of completed CG iterations
__global__ void uz_finish_kernel(const int *__restrict__ ctl, int *__restrict__ iters_done) { *iters_done = ctl[0] == 0 ? 1 : ctl[4]; }

} // namespace admmb200
