// kernels.cuh -- the CUDA kernels of one ADMM-elastic time step on sm_100a.
//
// Path (reference: admm::Solver::step, src/Solver.cpp:35-110):
//   step_begin_kernel      x_bar = x + dt v, M x_bar, curr_x = x_bar                 (:57-67)
//   tet_local_kernel       EnergyTerm::update for every tet: F = D_i x, z = prox(F+u),
//                          u += F - z  (src/EnergyTerm.hpp:130-140) fused with the element's
//                          share of dt^2 D^T W^2 (z-u)  (src/Solver.cpp:98)
//   tri_local_kernel       the same for TriEnergyTerm
//   pin_local_kernel       the same for SpringPin
//   assemble_kernel        b = M x_bar + sum of the per-corner shares (per-vertex segmented sum)
//   mcgs_kernel            NodalMultiColorGS::solve (src/NodalMultiColorGS.hpp:60-146), persistent,
//                          all sweeps x colours in one cooperative launch
//   step_end_kernel        v = (curr_x - x)/dt, x = curr_x                             (:105-106)
//
// Data layout (all in HBM, allocated once):
//   nodes      double4 per node (xyz + pad): x, v, curr_x, M x_bar, b, masses
//   tets       int4 idx[e]; E dminv[9][n_pad] (SoA, entry 3c+r); E u[9][n_pad]; E wdt2[n_pad];
//              E4 f[4e+c] = corner c's share of dt^2 D^T W^2 (z-u)   (E = float or double)
//   assembly   CSR node -> slots into f
//   MCGS       per colour sliced-ELL of the scalar matrix L (A = L (x) I3 + M), T lanes per node
#pragma once
#include "prox.cuh"
#include <cooperative_groups.h>

namespace admmb200 {

template <typename E> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; static __device__ __forceinline__ float4 make(float a, float b, float c) { return make_float4(a, b, c, 0.f); } };
template <> struct Vec4<double> { typedef double4 type; static __device__ __forceinline__ double4 make(double a, double b, double c) { return make_double4(a, b, c, 0.0); } };

__device__ __forceinline__ double4 ld_node(const double4 *p) {
	// 2 x 16-byte read-only loads
	const double2 *q = reinterpret_cast<const double2 *>(p);
	double2 a = __ldg(q), b = __ldg(q + 1);
	return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ double4 ld_node_cg(const double4 *p) {
	// L2-coherent loads: other SMs write these values inside the same kernel
	const double2 *q = reinterpret_cast<const double2 *>(p);
	double2 a = __ldcg(q), b = __ldcg(q + 1);
	return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_node(double4 *p, double x, double y, double z) {
	double2 *q = reinterpret_cast<double2 *>(p);
	q[0] = make_double2(x, y);
	q[1] = make_double2(z, 0.0);
}

// ---------------------------------------------------------------------------------------------
// step begin / end
// ---------------------------------------------------------------------------------------------
__global__ void step_begin_kernel(int n, double dt, double gravity, const double4 *__restrict__ x, double4 *__restrict__ v,
	const double4 *__restrict__ m, double4 *__restrict__ mxbar, double4 *__restrict__ cx)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double4 xi = x[i], vi = v[i], mi = m[i];
	if (fabs(gravity) > 0) { vi.y += dt * gravity; v[i] = vi; }
	double bx = xi.x + dt * vi.x, by = xi.y + dt * vi.y, bz = xi.z + dt * vi.z;
	st_node(&cx[i], bx, by, bz);
	st_node(&mxbar[i], mi.x * bx, mi.y * by, mi.z * bz);
}

// WindForce::project (src/ExplicitForce.cpp:47-104): a velocity kick per triangle, -alpha_n area v_n |v_n| n x 0.33 dt added to
// its three nodes, before gravity and x_bar (src/Solver.cpp:53-54).  The reference forms the kicks in an `omp parallel for`
// while other threads are already adding theirs to the same velocities: its result depends on the thread count.  Here every
// kick is formed from the velocities BEFORE the call (wind_tri_kernel), then each node adds the kicks of its triangles in a
// fixed order (wind_node_kernel): order independent and bit-reproducible.  oracle/admm_oracle.c: oracle_wind_project states
// both readings; they agree when the explicit drag is not stiff (alpha_n area |v| dt << 1).
__global__ void wind_tri_kernel(int n_tris, const int *__restrict__ tris, double dx, double dy, double dz, double dt,
	const double4 *__restrict__ x, const double4 *__restrict__ v, double4 *__restrict__ kick)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_tris) return;
	const int i0 = tris[3 * t], i1 = tris[3 * t + 1], i2 = tris[3 * t + 2];
	const double4 p0 = x[i0], p1 = x[i1], p2 = x[i2], v0 = v[i0], v1 = v[i1], v2 = v[i2];
	const double rx = (v0.x + v1.x + v2.x) / 3.0 - dx, ry = (v0.y + v1.y + v2.y) / 3.0 - dy, rz = (v0.z + v1.z + v2.z) / 3.0 - dz;
	const double ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z, bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
	double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
	const double len = sqrt(nx * nx + ny * ny + nz * nz), area = 0.5 * len;
	if (len > 0) { nx /= len; ny /= len; nz /= len; }
	const double vn = nx * rx + ny * ry + nz * rz;
	const double c = -1000.0 * area * vn * fabs(vn);
	st_node(&kick[t], c * nx * 0.33 * dt, c * ny * 0.33 * dt, c * nz * 0.33 * dt);
}
__global__ void wind_node_kernel(int n_touched, const int *__restrict__ nodes, const int *__restrict__ inc_ptr, const int *__restrict__ inc_tri,
	const double4 *__restrict__ kick, double4 *__restrict__ v)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_touched) return;
	const int i = nodes[k];
	double4 vi = v[i];
	for (int q = inc_ptr[k]; q < inc_ptr[k + 1]; ++q) { const double4 f = kick[inc_tri[q]]; vi.x += f.x; vi.y += f.y; vi.z += f.z; }
	st_node(&v[i], vi.x, vi.y, vi.z);
}

__global__ void step_end_kernel(int n, double dt, double4 *__restrict__ x, double4 *__restrict__ v, const double4 *__restrict__ cx)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double4 xi = x[i], ci = cx[i];
	double inv = 1.0 / dt;
	st_node(&v[i], (ci.x - xi.x) * inv, (ci.y - xi.y) * inv, (ci.z - xi.z) * inv);
	st_node(&x[i], ci.x, ci.y, ci.z);
}

// ---------------------------------------------------------------------------------------------
// local step: tets
// ---------------------------------------------------------------------------------------------
template <typename E>
struct TetBatch {
	int n, n_pad;             // elements, SoA pitch
	const int4 *idx;          // [n]
	const E *dminv;           // [9][n_pad], entry 3c+r = edges_inv(c,r)
	const E *wdt2;            // [n_pad]   dt^2 w^2
	E *u;                     // [9][n_pad]
	E *z;                     // [9][n_pad] or NULL
	E *q;                     // [4][n_pad] V of the element's last SVD as a quaternion (warm start, see svd3_signed), or NULL
	typename Vec4<E>::type *f; // [4n]
	Material<E> mat;
	int *defer_count;         // queue of degenerate elements (see prox.cuh, PROX_FAST / PROX_REFERENCE)
	int *defer_list;          // [n]
	int *defer_done;          // blocks of the consumer kernel that have finished (the last one empties the queue)
};

// The streamed part of one element: indices and the SoA columns (everything that comes from DRAM).
// SVD warm start (prox.cuh, svd3_signed): compiled out by default -- measured slower on the 1M-tet beam (61.8 vs 59.7 us)
#ifndef ADMMB200_SVD_WARMSTART
#define ADMMB200_SVD_WARMSTART 0
#endif
template <typename E> struct TetStream { int4 id; E bi[9], u[9], w; E q[ADMMB200_SVD_WARMSTART ? 4 : 1]; };

template <typename E>
__device__ __forceinline__ void tet_load(const TetBatch<E> &tb, int e, TetStream<E> &t)
{
	const int np = tb.n_pad;
	// Load order matters: the vertex gathers depend on the index load, the SoA columns do not -- issue
	// index first, then all independent columns, and only then touch the indices.
	t.id = __ldg(&tb.idx[e]);
#pragma unroll
	for (int k = 0; k < 9; ++k) t.bi[k] = __ldg(&tb.dminv[(size_t)k * np + e]);
#pragma unroll
	for (int k = 0; k < 9; ++k) t.u[k] = tb.u[(size_t)k * np + e];
	t.w = __ldg(&tb.wdt2[e]);
	if (ADMMB200_SVD_WARMSTART && tb.q) {
#pragma unroll
		for (int k = 0; k < (ADMMB200_SVD_WARMSTART ? 4 : 1); ++k) t.q[k] = tb.q[(size_t)k * np + e];
	} else t.q[0] = E(0);
}

// One element of the local step.  MODE = PROX_FAST: the hot kernel; a degenerate element is queued and
// left untouched.  MODE = PROX_REFERENCE: the queue's consumer redoes such an element from scratch.
template <typename E, int MODEL, bool STORE_Z, int MODE>
__device__ __forceinline__ void tet_compute(const TetBatch<E> &tb, const double4 *__restrict__ cx, int e, TetStream<E> &t)
{
	const int np = tb.n_pad;
	const int4 id = t.id;
	E (&bi)[9] = t.bi; E (&u)[9] = t.u;
	E *q = t.q;
	const E w = t.w;
	double4 p0 = ld_node(&cx[id.x]), p1 = ld_node(&cx[id.y]), p2 = ld_node(&cx[id.z]), p3 = ld_node(&cx[id.w]);
	// Ds = [x1-x0, x2-x0, x3-x0], differences in fp64 (positions are ~metres, edges ~centimetres)
	E ds[9] = {E(p1.x - p0.x), E(p1.y - p0.y), E(p1.z - p0.z), E(p2.x - p0.x), E(p2.y - p0.y), E(p2.z - p0.z), E(p3.x - p0.x), E(p3.y - p0.y), E(p3.z - p0.z)};
	// F = Ds * Binv, column-major F[3r+j] = sum_c Ds(j,c) Binv(c,r)   (D_i x, src/TetEnergyTerm.cpp:50-71)
	// zin = D_i x + u_i is all that has to survive the prox: u_new = u + D_i x - z = zin - z (EnergyTerm::update,
	// src/EnergyTerm.hpp:130-140), so F, u and Dm^-1 are dead across the SVD + Newton iteration (register pressure
	// there decides the occupancy, and the occupancy the speed); Dm^-1 is read again afterwards (L1/L2 hit).
	E zin[9], z[9];
#pragma unroll
	for (int r = 0; r < 3; ++r)
#pragma unroll
		for (int j = 0; j < 3; ++j) {
			zin[3 * r + j] = (ds[j] * bi[r] + ds[3 + j] * bi[3 + r] + ds[6 + j] * bi[6 + r]) + u[3 * r + j];
			z[3 * r + j] = zin[3 * r + j];
		}
	if (prox_tet_mode<E, MODEL, MODE>(tb.mat, z, (ADMMB200_SVD_WARMSTART && tb.q) ? q : nullptr)) {
		tb.defer_list[atomicAdd(tb.defer_count, 1)] = e; // MODE == PROX_FAST only
		return;
	}
	if (ADMMB200_SVD_WARMSTART && tb.q) {
#pragma unroll
		for (int k = 0; k < (ADMMB200_SVD_WARMSTART ? 4 : 1); ++k) tb.q[(size_t)k * np + e] = q[k];
	}
	// u += Dx - z ; y = z - u_new
	E y[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) {
		const E un = zin[k] - z[k];
		tb.u[(size_t)k * np + e] = un;
		if (STORE_Z) tb.z[(size_t)k * np + e] = z[k];
		y[k] = z[k] - un;
	}
	E b2[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) b2[k] = __ldg(&tb.dminv[(size_t)k * np + e]);
	// corner shares of dt^2 D^T W^2 y: corner c>=1: wdt2 * sum_r Binv(c-1,r) y[:,r]; corner 0: minus their sum
	E f1[3], f2[3], f3[3];
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		f1[j] = w * (b2[0] * y[j] + b2[1] * y[3 + j] + b2[2] * y[6 + j]);
		f2[j] = w * (b2[3] * y[j] + b2[4] * y[3 + j] + b2[5] * y[6 + j]);
		f3[j] = w * (b2[6] * y[j] + b2[7] * y[3 + j] + b2[8] * y[6 + j]);
	}
	typename Vec4<E>::type *f = tb.f + (size_t)4 * e;
	f[0] = Vec4<E>::make(-(f1[0] + f2[0] + f3[0]), -(f1[1] + f2[1] + f3[1]), -(f1[2] + f2[2] + f3[2]));
	f[1] = Vec4<E>::make(f1[0], f1[1], f1[2]);
	f[2] = Vec4<E>::make(f2[0], f2[1], f2[2]);
	f[3] = Vec4<E>::make(f3[0], f3[1], f3[2]);
}

template <typename E, int MODEL, bool STORE_Z, int MODE>
__device__ __forceinline__ void tet_element(const TetBatch<E> &tb, const double4 *__restrict__ cx, int e)
{
	TetStream<E> t;
	tet_load(tb, e, t);
	tet_compute<E, MODEL, STORE_Z, MODE>(tb, cx, e, t);
}

// MINB = resident blocks per SM asked of the compiler (register budget 65536 / (128 MINB))
template <typename E, int MODEL, bool STORE_Z, int MINB>
__global__ void __launch_bounds__(128, MINB) tet_local_kernel(TetBatch<E> tb, const double4 *__restrict__ cx)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= tb.n) return;
	tet_element<E, MODEL, STORE_Z, PROX_FAST>(tb, cx, e);
}

// Consumer of the queue of degenerate elements (usually empty: the launch costs a few microseconds).
template <typename E, int MODEL, bool STORE_Z>
__global__ void __launch_bounds__(128) tet_local_deferred_kernel(TetBatch<E> tb, const double4 *__restrict__ cx)
{
	const int count = *tb.defer_count;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x)
		tet_element<E, MODEL, STORE_Z, PROX_REFERENCE>(tb, cx, tb.defer_list[k]);
	// the last block to finish empties the queue for the next local step (every block has read `count` by then),
	// so the hot kernel needs no memset in front of it
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(tb.defer_done, 1) == (int)gridDim.x - 1) { *tb.defer_count = 0; *tb.defer_done = 0; }
	}
}

// prox alone on raw deformation gradients (parity tests / micro-benchmarks): zio is [9][n_pad] SoA
template <typename E, int MODEL>
__global__ void __launch_bounds__(128) tet_prox_only_kernel(int n, int n_pad, E *zio, Material<E> mat, int *defer_count, int *defer_list)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n) return;
	E z[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) z[k] = zio[(size_t)k * n_pad + e];
	if (prox_tet_mode<E, MODEL, PROX_FAST>(mat, z)) { defer_list[atomicAdd(defer_count, 1)] = e; return; }
#pragma unroll
	for (int k = 0; k < 9; ++k) zio[(size_t)k * n_pad + e] = z[k];
}
template <typename E, int MODEL>
__global__ void __launch_bounds__(128) tet_prox_only_deferred_kernel(int n_pad, E *zio, Material<E> mat, const int *defer_count, const int *defer_list)
{
	const int count = *defer_count;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
		const int e = defer_list[k];
		E z[9];
#pragma unroll
		for (int j = 0; j < 9; ++j) z[j] = zio[(size_t)j * n_pad + e];
		prox_tet_mode<E, MODEL, PROX_REFERENCE>(mat, z);
#pragma unroll
		for (int j = 0; j < 9; ++j) zio[(size_t)j * n_pad + e] = z[j];
	}
}

// ---------------------------------------------------------------------------------------------
// local step: triangles
// ---------------------------------------------------------------------------------------------
template <typename E>
struct TriBatch {
	int n, n_pad;
	const int4 *idx;   // [n] (w unused)
	const E *rest;     // [4][n_pad], entry 2c+r = rest_pose(c,r)
	const E *wdt2;     // [n_pad]
	E *u;              // [6][n_pad]
	E *z;              // [6][n_pad] or NULL
	typename Vec4<E>::type *f; // [3n]
	E limit_min, limit_max;
};

template <typename E, bool STORE_Z>
__global__ void __launch_bounds__(128) tri_local_kernel(TriBatch<E> tb, const double4 *__restrict__ cx)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= tb.n) return;
	const int np = tb.n_pad;
	int4 id = __ldg(&tb.idx[e]);
	double4 p0 = ld_node(&cx[id.x]), p1 = ld_node(&cx[id.y]), p2 = ld_node(&cx[id.z]);
	E ds[6] = {E(p1.x - p0.x), E(p1.y - p0.y), E(p1.z - p0.z), E(p2.x - p0.x), E(p2.y - p0.y), E(p2.z - p0.z)};
	E rp[4], u[6];
#pragma unroll
	for (int k = 0; k < 4; ++k) rp[k] = tb.rest[(size_t)k * np + e];
#pragma unroll
	for (int k = 0; k < 6; ++k) u[k] = tb.u[(size_t)k * np + e];
	// F (3x2) = [x1-x0, x2-x0] * rest_pose; rows of D: i for column 0, 3+i for column 1 (src/TriEnergyTerm.cpp:54-70)
	E F[6], z[6];
#pragma unroll
	for (int r = 0; r < 2; ++r)
#pragma unroll
		for (int j = 0; j < 3; ++j) {
			F[3 * r + j] = ds[j] * rp[r] + ds[3 + j] * rp[2 + r];
			z[3 * r + j] = F[3 * r + j] + u[3 * r + j];
		}
	prox_tri<E>(tb.limit_min, tb.limit_max, z);
	E y[6];
#pragma unroll
	for (int k = 0; k < 6; ++k) {
		E un = u[k] + (F[k] - z[k]);
		tb.u[(size_t)k * np + e] = un;
		if (STORE_Z) tb.z[(size_t)k * np + e] = z[k];
		y[k] = z[k] - un;
	}
	E w = tb.wdt2[e];
	E f1[3], f2[3];
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		f1[j] = w * (rp[0] * y[j] + rp[1] * y[3 + j]);
		f2[j] = w * (rp[2] * y[j] + rp[3] * y[3 + j]);
	}
	typename Vec4<E>::type *f = tb.f + (size_t)3 * e;
	f[0] = Vec4<E>::make(-(f1[0] + f2[0]), -(f1[1] + f2[1]), -(f1[2] + f2[2]));
	f[1] = Vec4<E>::make(f1[0], f1[1], f1[2]);
	f[2] = Vec4<E>::make(f2[0], f2[1], f2[2]);
}

template <typename E>
__global__ void __launch_bounds__(128) tri_prox_only_kernel(int n, int n_pad, E *zio, E limit_min, E limit_max)
{
	int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n) return;
	E z[6];
#pragma unroll
	for (int k = 0; k < 6; ++k) z[k] = zio[(size_t)k * n_pad + e];
	prox_tri<E>(limit_min, limit_max, z);
#pragma unroll
	for (int k = 0; k < 6; ++k) zio[(size_t)k * n_pad + e] = z[k];
}

// ---------------------------------------------------------------------------------------------
// local step: SpringPin (src/SpringEnergyTerm.hpp:31-73).  Always fp64: a handful of elements.
// ---------------------------------------------------------------------------------------------
struct PinBatch {
	int n;
	const int *idx;        // [n]
	const double *pos;     // [3n]
	const unsigned char *active; // [n]
	const double *wdt2;    // [n]
	double *u;             // [3n]
	double *z;             // [3n]
	double4 *f;            // [n]
};

__global__ void pin_local_kernel(PinBatch pb, const double4 *__restrict__ cx)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= pb.n) return;
	double4 p = cx[pb.idx[i]];
	double dix[3] = {p.x, p.y, p.z};
	double out[3];
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		double u = pb.u[3 * i + j];
		double z = dix[j] + u;
		if (pb.active[i]) z = pb.pos[3 * i + j];
		double un = u + (dix[j] - z);
		pb.u[3 * i + j] = un;
		pb.z[3 * i + j] = z;
		out[j] = pb.wdt2[i] * (z - un);
	}
	pb.f[i] = make_double4(out[0], out[1], out[2], 0.0);
}

// ---------------------------------------------------------------------------------------------
// global step 1: b = M x_bar + dt^2 D^T W^2 (z-u)   (src/Solver.cpp:98), per-vertex segmented sum
// over the corner shares written by the local kernels, always in the same order (no float atomics).
// Incidence: sliced and transposed per warp -- the 32 vertices of a warp share rows [inc_off[w], inc_off[w+1]),
// row j holds the j-th corner of each of them (inc_slot[row * 32 + lane]), so the index loads are coalesced and
// independent of each other.  Entry >= 0: slot in the element-precision share array; < 0: fp64 pin share
// (low 31 bits); ADMMB200_NO_SLOT: padding.
// ---------------------------------------------------------------------------------------------
#define ADMMB200_NO_SLOT 0x7fffffff
template <typename E>
__global__ void __launch_bounds__(256) assemble_kernel(int n, const int *__restrict__ inc_off, const int *__restrict__ inc_slot,
	const typename Vec4<E>::type *__restrict__ f, const double4 *__restrict__ fpin, const double4 *__restrict__ mxbar, double4 *__restrict__ b)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int w = i >> 5, lane = threadIdx.x & 31;
	if (w * 32 >= n) return;
	const int r0 = __ldg(&inc_off[w]), r1 = __ldg(&inc_off[w + 1]);
	double sx = 0, sy = 0, sz = 0;
	const int *sl = inc_slot + (size_t)r0 * 32 + lane;
	const int nr = r1 - r0;
	// batches of 8 rows: the 8 slot indices first (coalesced, independent), then the 8 scattered 16-byte share loads
	// together -- the kernel is bound by the latency of those gathers, so what counts is how many are in flight
	for (int r = 0; r < nr; r += 8) {
		int sidx[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) sidx[j] = (r + j < nr) ? __ldg(sl + (size_t)(r + j) * 32) : ADMMB200_NO_SLOT;
		typename Vec4<E>::type v[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			v[j].x = 0; v[j].y = 0; v[j].z = 0; v[j].w = 0;
			if (sidx[j] >= 0 && sidx[j] != ADMMB200_NO_SLOT) v[j] = f[sidx[j]];
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			sx += double(v[j].x); sy += double(v[j].y); sz += double(v[j].z);
			if (sidx[j] < 0) { const double4 p = fpin[sidx[j] & 0x7fffffff]; sx += p.x; sy += p.y; sz += p.z; } // fp64 pin share (rare)
		}
	}
	if (i >= n) return;
	const double4 acc = mxbar[i];
	st_node(&b[i], acc.x + sx, acc.y + sy, acc.z + sz);
}

// ---------------------------------------------------------------------------------------------
// global step 2: nodal multi-colour Gauss-Seidel, persistent cooperative kernel.
// ---------------------------------------------------------------------------------------------
struct Obstacle { int kind; double p[4]; };
#define ADMMB200_MAX_OBSTACLES 8

struct McgsParams {
	int n_nodes;
	int n_colors;
	int iters;                  // max_iters
	double omega;
	double tol2;                // m_tol^2, <= 0: no residual test
	const int *color_first_slice; // [n_colors+1]
	const int *slice_ptr;       // [n_slices+1], offset (in 32-entry rows) into ell_col / ell_val
	const int *slice_node;      // [n_slices * (32/T)] node of each group, -1 = padding
	const int *ell_col;         // [rows*32]
	const double *ell_val;      // [rows*32]
	const double *diag;         // [3*n_nodes]  a_ii per component (L_ii + m)
	const int *pin_slot;        // [n_nodes] -1 or index into pin_pos
	const double *pin_pos;      // [3*n_pins]
	int has_pins;
	int n_obstacles;
	const Obstacle *obs;        // [n_obstacles] in global memory
	double4 *x;                 // curr_x, in/out
	const double4 *b;
	unsigned int *barrier;      // zeroed before launch
	double *resid;              // [iters+1], zeroed before launch: [0] = |b|^2, [1+it] = |b-Ax|^2 after sweep it
	double *resid_lb;           // [iters], zeroed before launch: lower bound of |b-Ax|^2 after sweep it
	int *iters_done;            // return value of solve()
};

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
	unsigned int v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
	asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u32(unsigned int *p, unsigned int v) {
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Grid-wide barrier of a cooperative (co-resident) launch.  Polling uses relaxed loads (an acquire load
// would invalidate L1 on every poll); one acq_rel fence on each side gives the ordering.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &target, unsigned int n_blocks)
{
	__syncthreads();
	if (threadIdx.x == 0) {
		target += n_blocks;
		fence_acq_rel_gpu();
		atomicAdd(counter, 1u);
		while (ld_relaxed_u32(counter) < target) { }
		fence_acq_rel_gpu();
	}
	__syncthreads();
}

// Creates the ortho projection of NodalMultiColorGS::orthoG (src/NodalMultiColorGS.hpp:171-177)
// and applies constrained_segment_update's plane solve (:249-259): G G^T (x_gs - p) + p.
__device__ __forceinline__ void plane_project(const double *n, const double *p, const double *xgs, double *out)
{
	double nn0 = n[0] > 0.999 ? 0.0 : 1.0, nn1 = 0.0, nn2 = n[0] > 0.999 ? 1.0 : 0.0;
	double u0 = nn1 * n[2] - nn2 * n[1], u1 = nn2 * n[0] - nn0 * n[2], u2 = nn0 * n[1] - nn1 * n[0];
	double iu = 1.0 / sqrt(u0 * u0 + u1 * u1 + u2 * u2); u0 *= iu; u1 *= iu; u2 *= iu;
	double v0 = n[1] * u2 - n[2] * u1, v1 = n[2] * u0 - n[0] * u2, v2 = n[0] * u1 - n[1] * u0;
	double iv = 1.0 / sqrt(v0 * v0 + v1 * v1 + v2 * v2); v0 *= iv; v1 *= iv; v2 *= iv;
	double d0 = xgs[0] - p[0], d1 = xgs[1] - p[1], d2 = xgs[2] - p[2];
	double t0 = u0 * d0 + u1 * d1 + u2 * d2, t1 = v0 * d0 + v1 * d1 + v2 * d2;
	out[0] = (u0 * t0 + v0 * t1) + p[0];
	out[1] = (u1 * t0 + v1 * t1) + p[1];
	out[2] = (u2 * t0 + v2 * t1) + p[2];
}

// Collider::detect_passive (src/Collider.hpp:137-150) for Floor / Sphere (src/PassiveObject.hpp:32-64)
__device__ __forceinline__ bool detect_passive(const Obstacle *obs, int n_obstacles, const double *x, double *n, double *p)
{
	double dx = 1.7976931348623157e308;
	for (int j = 0; j < n_obstacles; ++j) {
		const Obstacle o = obs[j];
		if (o.kind == 0) {
			double d = x[1] - o.p[0];
			if (!(d > dx)) { dx = d; p[0] = x[0]; p[1] = o.p[0]; p[2] = x[2]; n[0] = 0; n[1] = 1; n[2] = 0; }
		} else {
			double r0 = x[0] - o.p[0], r1 = x[1] - o.p[1], r2 = x[2] - o.p[2];
			double len = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
			double d = len - o.p[3];
			if (!(d > dx)) {
				dx = d; r0 /= len; r1 /= len; r2 /= len;
				p[0] = o.p[0] + r0 * o.p[3]; p[1] = o.p[1] + r1 * o.p[3]; p[2] = o.p[2] + r2 * o.p[3];
				n[0] = r0; n[1] = r1; n[2] = r2;
			}
		}
		if (dx < 0) return true;
	}
	return false;
}

#define ADMMB200_MCGS_THREADS 768

// One gather over the sliced-ELL rows [r0, r1) of a slice: every lane accumulates its entries, then the
// T lanes of a node are summed with shuffles.  x is read with L2-coherent loads (other SMs wrote it in
// the previous colour pass); col/val/b/diag never change during the kernel and take the read-only path.
template <int T>
__device__ __forceinline__ void mcgs_row_gather(const McgsParams &P, int r0, int r1, int lane, double &sx, double &sy, double &sz)
{
	sx = 0; sy = 0; sz = 0;
	const int *col = P.ell_col + (size_t)r0 * 32 + lane;
	const double *val = P.ell_val + (size_t)r0 * 32 + lane;
	const int n = r1 - r0;
#pragma unroll 4
	for (int r = 0; r < n; ++r) {
		int c = __ldg(col + (size_t)r * 32);
		double a = __ldg(val + (size_t)r * 32);
		double4 xc = ld_node_cg(&P.x[c]);
		sx += a * xc.x; sy += a * xc.y; sz += a * xc.z;
	}
#pragma unroll
	for (int o = 1; o < T; o <<= 1) {
		sx += __shfl_xor_sync(0xffffffffu, sx, o);
		sy += __shfl_xor_sync(0xffffffffu, sy, o);
		sz += __shfl_xor_sync(0xffffffffu, sz, o);
	}
}

// Passive obstacles inside the sweep (src/NodalMultiColorGS.hpp:124-127), out of line: rare.
__device__ __noinline__ bool mcgs_collide(const Obstacle *obs, int n_obstacles, const double *gs, double *nx)
{
	double nrm[3], pt[3];
	if (!detect_passive(obs, n_obstacles, nx, nrm, pt)) return false;
	double out[3];
	plane_project(nrm, pt, gs, out); // constrained_segment_update (:218-262): no over-relaxation
	nx[0] = out[0]; nx[1] = out[1]; nx[2] = out[2];
	return true;
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0;
	if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
	return s; // valid in thread 0
}

// The convergence test of the reference, |b - A x|^2 / |b|^2 < tol^2 after every sweep
// (src/NodalMultiColorGS.hpp:136-139), costs a full extra pass over the matrix and in practice never
// fires.  It is evaluated lazily without changing its outcome: a node i of the LAST
// colour has, right after its SOR update, the residual row r_i = a_ii (1/omega - 1) dx_i exactly (no
// later colour touches its neighbours), and the sum of r_i^2 over any subset of rows is a lower bound of
// |b - A x|^2.  Only when that bound is below 4x the threshold is the exact residual computed.
template <int T>
__global__ void __launch_bounds__(ADMMB200_MCGS_THREADS, 1) mcgs_kernel(McgsParams P)
{
	constexpr int G = 32 / T; // nodes per slice
	const int lane = threadIdx.x & 31;
	const int sub = lane % T;
	const int grp = lane / T;
	const int warps_per_block = blockDim.x >> 5;
	const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
	const int n_warps = gridDim.x * warps_per_block;
	unsigned int bar_target = 0;
	__shared__ double red[32];
	const bool check = P.tol2 > 0.0;
	const double omega = P.omega, one_m_omega = 1.0 - P.omega, lb_scale = 1.0 / P.omega - 1.0;

	if (check) {
		// b_norm = |b|^2 (src/NodalMultiColorGS.hpp:92)
		double acc = 0;
		for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_nodes; i += gridDim.x * blockDim.x) {
			double4 bi = ld_node(&P.b[i]);
			acc += bi.x * bi.x + bi.y * bi.y + bi.z * bi.z;
		}
		double s = block_sum(acc, red);
		if (threadIdx.x == 0) atomicAdd(&P.resid[0], s);
	}

	int it = 0;
	for (; it < P.iters; ++it) {
		double lb = 0;
		for (int color = 0; color < P.n_colors; ++color) {
			const int s0 = __ldg(&P.color_first_slice[color]), s1 = __ldg(&P.color_first_slice[color + 1]);
			const bool last = check && (color == P.n_colors - 1);
			for (int sl = s0 + warp_global; sl < s1; sl += n_warps) {
				const int node = __ldg(&P.slice_node[sl * G + grp]);
				const int r0 = __ldg(&P.slice_ptr[sl]), r1 = __ldg(&P.slice_ptr[sl + 1]);
				const bool owner = (sub == 0 && node >= 0);
				// the node's own data does not depend on the gather: issue these loads first
				double4 bi = make_double4(0, 0, 0, 0), xi = bi;
				double a0 = 1, a1 = 1, a2 = 1;
				int ps = -1;
				if (owner) {
					bi = ld_node(&P.b[node]);
					xi = ld_node_cg(&P.x[node]);
					a0 = __ldg(&P.diag[3 * node]); a1 = __ldg(&P.diag[3 * node + 1]); a2 = __ldg(&P.diag[3 * node + 2]);
					if (P.has_pins) ps = __ldg(&P.pin_slot[node]);
				}
				double sx, sy, sz;
				mcgs_row_gather<T>(P, r0, r1, lane, sx, sy, sz);
				if (owner) {
					if (ps >= 0) {
						st_node(&P.x[node], P.pin_pos[3 * ps], P.pin_pos[3 * ps + 1], P.pin_pos[3 * ps + 2]);
					} else {
						// segment_update (src/NodalMultiColorGS.hpp:180-215)
						double gs[3] = {(bi.x - sx) / a0, (bi.y - sy) / a1, (bi.z - sz) / a2};
						double nx[3] = {one_m_omega * xi.x + omega * gs[0], one_m_omega * xi.y + omega * gs[1], one_m_omega * xi.z + omega * gs[2]};
						bool hit = false;
						if (P.n_obstacles > 0) hit = mcgs_collide(P.obs, P.n_obstacles, gs, nx);
						st_node(&P.x[node], nx[0], nx[1], nx[2]);
						if (last && !hit) {
							double rx = a0 * lb_scale * (nx[0] - xi.x), ry = a1 * lb_scale * (nx[1] - xi.y), rz = a2 * lb_scale * (nx[2] - xi.z);
							lb += rx * rx + ry * ry + rz * rz;
						}
					}
				}
			}
			if (last) {
				double s = block_sum(lb, red);
				if (threadIdx.x == 0 && s > 0.0) atomicAdd(&P.resid_lb[it], s);
			}
			grid_barrier(P.barrier, bar_target, gridDim.x);
		}
		if (check) {
			const double b2 = __ldcg(&P.resid[0]);
			const double bound = __ldcg(&P.resid_lb[it]);
			if (!(bound >= 4.0 * P.tol2 * b2)) {
				// exact residual = b - A x (src/NodalMultiColorGS.hpp:136-139), every row including pinned ones
				double acc = 0;
				const int n_slices = P.color_first_slice[P.n_colors];
				for (int sl = warp_global; sl < n_slices; sl += n_warps) {
					const int node = __ldg(&P.slice_node[sl * G + grp]);
					const int r0 = __ldg(&P.slice_ptr[sl]), r1 = __ldg(&P.slice_ptr[sl + 1]);
					double sx, sy, sz;
					mcgs_row_gather<T>(P, r0, r1, lane, sx, sy, sz);
					if (sub == 0 && node >= 0) {
						double4 bi = ld_node(&P.b[node]);
						double4 xi = ld_node_cg(&P.x[node]);
						double rx = bi.x - (sx + P.diag[3 * node] * xi.x);
						double ry = bi.y - (sy + P.diag[3 * node + 1] * xi.y);
						double rz = bi.z - (sz + P.diag[3 * node + 2] * xi.z);
						acc += rx * rx + ry * ry + rz * rz;
					}
				}
				double s = block_sum(acc, red);
				if (threadIdx.x == 0) atomicAdd(&P.resid[1 + it], s);
				grid_barrier(P.barrier, bar_target, gridDim.x);
				double r2 = __ldcg(&P.resid[1 + it]);
				if (r2 / b2 < P.tol2) break;
			}
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) *P.iters_done = it;
}

// ---------------------------------------------------------------------------------------------
// helpers: AoS <-> padded / SoA conversions done on the device so host copies stay contiguous
// ---------------------------------------------------------------------------------------------
__global__ void pack3_to4_kernel(int n, const double *__restrict__ in3, double4 *__restrict__ out4)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	st_node(&out4[i], in3[3 * i], in3[3 * i + 1], in3[3 * i + 2]);
}
__global__ void unpack4_to3_kernel(int n, const double4 *__restrict__ in4, double *__restrict__ out3)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double4 v = in4[i];
	out3[3 * i] = v.x; out3[3 * i + 1] = v.y; out3[3 * i + 2] = v.z;
}

} // namespace admmb200
