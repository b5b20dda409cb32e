// prox.cuh -- per-element proximal operators of the ADMM local step, in registers.
//
// One thread owns one element.  Everything is templated on the element scalar T
// (float = production path, double = validation path).
//
// What is computed (parity targets in the reference, mattoverby/admm-elastic @ c6c09a3):
//   svd3_signed       signed_svd            src/FastSVD.hpp:43-68  (a stub around Eigen::JacobiSVD;
//                                           parity target is U*diag(f(S))*V^T, not U and V themselves)
//   prox_tet_linear   TetEnergyTerm::prox   src/TetEnergyTerm.cpp:73-92
//   prox_tet_hyper    HyperElasticTet::prox src/TetEnergyTerm.cpp:114-136 with the energies of
//                     NHProx :173-204, StVKProx :210-237, SplineProx :243-265 + src/XuSpline.hpp:48-94
//   prox_tri          TriEnergyTerm::prox   src/TriEnergyTerm.cpp:73-101
//
// The reference minimises the 3-variable prox objective with L-BFGS + cubic backtracking
// (deps/mcloptlib LBFGS.hpp:52-152) to |grad|<1e-6 or |dx|<1e-6; here it is a safeguarded Newton
// iteration with the closed-form 3x3 Hessian, which converges to the same minimiser in 3-6
// iterations (SURVEY.md 0.2, 7 "hard parts" 1-2).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// Every function here is pure register arithmetic.  ADMMB200_FN is __device__ in the product build;
// tests/tools/prox_host.cu recompiles the same source for the host (ADMMB200_HOST_SHIM) so that
// parity problems can be studied without a GPU.  The product library never defines that macro.
#ifdef ADMMB200_HOST_SHIM
#define ADMMB200_FN __host__ __device__ __forceinline__
#define ADMMB200_SLOWPATH __host__ __device__ __noinline__
#else
#define ADMMB200_FN __device__ __forceinline__
#define ADMMB200_SLOWPATH __device__ __noinline__
#endif

namespace admmb200 {

enum TetModel { TET_LINEAR = 0, TET_NEOHOOKEAN = 1, TET_STVK = 2, TET_SPLINE_NH = 3, TET_SPLINE_STVK = 4, TET_SPLINE_COROT = 5 };

template <typename T> struct Num;
template <> struct Num<float> {
	static ADMMB200_FN float eps() { return 1.1920929e-7f; }
	static ADMMB200_FN float tiny() { return 1e-30f; }
	static ADMMB200_FN float rsqrt(float x) { return ::rsqrtf(x); }
	static ADMMB200_FN float sqrt(float x) { return ::sqrtf(x); }
	static ADMMB200_FN float log(float x) { return ::logf(x); }
	// approximate (<= 2 ulp) quotient / reciprocal for quantities that only steer an iteration (Jacobi
	// angles, Newton directions): one MUFU.RCP instead of the IEEE sequence with its slow-path branch
#if defined(__CUDA_ARCH__)
	static ADMMB200_FN float fdiv(float a, float b) { return __fdividef(a, b); }
	static ADMMB200_FN float frcp(float x) { return __fdividef(1.0f, x); }
#else
	static ADMMB200_FN float fdiv(float a, float b) { return a / b; }
	static ADMMB200_FN float frcp(float x) { return 1.0f / x; }
#endif
	static constexpr int jacobi_sweeps = 5;
	static constexpr int newton_iters = 24;
	static ADMMB200_FN float newton_tol() { return 2e-4f; } // quadratic convergence: the step after a 2e-4 step is < 1e-7
};
template <> struct Num<double> {
	static ADMMB200_FN double eps() { return 2.220446049250313e-16; }
	static ADMMB200_FN double tiny() { return 1e-290; }
	static ADMMB200_FN double rsqrt(double x) { return 1.0 / ::sqrt(x); }
	static ADMMB200_FN double sqrt(double x) { return ::sqrt(x); }
	static ADMMB200_FN double log(double x) { return ::log(x); }
	static ADMMB200_FN double fdiv(double a, double b) { return a / b; }
	static ADMMB200_FN double frcp(double x) { return 1.0 / x; }
	static constexpr int jacobi_sweeps = 8;
	static constexpr int newton_iters = 48;
	static ADMMB200_FN double newton_tol() { return 1e-7; }
};

// One Jacobi rotation in the (p,q) plane of a symmetric 3x3 matrix; r is the third index.
// A' = P^T A P, V' = V P with P a proper rotation, so det V stays +1.
//
// The angle (|phi| <= pi/4, tan 2phi = 2 apq / (aqq - app)) comes from the half-angle identities instead
// of the textbook t = sign(theta) / (|theta| + sqrt(theta^2 + 1)): with alpha = aqq - app, beta = 2 apq,
//     cos 2phi = |alpha| / hypot(alpha, beta),  c^2 = (1 + cos 2phi) / 2,  s = sign(alpha) beta / (2 c hypot)
// which is two reciprocal square roots (MUFU.RSQ on the device) and no division, no square root, no
// branch.  apq = 0 gives c = 1, s = 0 exactly, so a negligible off-diagonal needs no special case; only
// alpha = beta = 0 (hypot underflows) is forced to the identity.
template <typename T>
ADMMB200_FN void jacobi_rot(T &app, T &aqq, T &apq, T &arp, T &arq,
	T &v0p, T &v0q, T &v1p, T &v1q, T &v2p, T &v2q)
{
	const T alpha = aqq - app, beta = apq + apq;
	const T r2 = alpha * alpha + beta * beta;
	const bool ok = r2 > Num<T>::tiny();
	const T ir = Num<T>::rsqrt(ok ? r2 : T(1));
	const T c2 = ok ? T(0.5) + T(0.5) * fabs(alpha) * ir : T(1);
	const T ic = Num<T>::rsqrt(c2);
	const T c = c2 * ic;
	const T s = copysign(T(0.5) * ir * ic, alpha) * (ok ? beta : T(0));
	const T t = s * ic; // tan phi
	app -= t * apq;
	aqq += t * apq;
	apq = T(0);
	T nrp = c * arp - s * arq, nrq = s * arp + c * arq; arp = nrp; arq = nrq;
	T a, b;
	a = c * v0p - s * v0q; b = s * v0p + c * v0q; v0p = a; v0q = b;
	a = c * v1p - s * v1q; b = s * v1p + c * v1q; v1p = a; v1q = b;
	a = c * v2p - s * v2q; b = s * v2p + c * v2q; v2p = a; v2q = b;
}

// F (column-major, F[3c+r] = F(r,c)) = U diag(S) V^T with U, V in SO(3), S[0] >= S[1] >= |S[2]|,
// sign(S[2]) = sign(det F): the convention signed_svd (src/FastSVD.hpp:43-68) produces.
// U, V are column-major too.
//
// Warm start: q (in/out, may be null) is V of the previous call for this element as a quaternion (w,x,y,z),
// not necessarily normalised; all zeros = no guess.  Between two ADMM iterations F + u moves little, so
// V0^T (F^T F) V0 is already nearly diagonal and the Jacobi iteration needs one or two sweeps instead of
// four.  A quaternion (normalised on load) cannot drift away from a rotation, however often it is reused.
template <typename T>
ADMMB200_FN void svd3_signed(const T *F, T *S, T *U, T *V, T *q = nullptr)
{
	// V(r,c): v{r}{c}
	T v00 = 1, v01 = 0, v02 = 0, v10 = 0, v11 = 1, v12 = 0, v20 = 0, v21 = 0, v22 = 1;
	T c00, c01, c02, c11, c12, c22;
	if (q) {
		const T qq = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
		const bool have = qq > Num<T>::tiny();
		const T n = Num<T>::rsqrt(have ? qq : T(1));
		const T w = have ? q[0] * n : T(1), x = q[1] * n, y = q[2] * n, z = q[3] * n;
		v00 = T(1) - T(2) * (y * y + z * z); v01 = T(2) * (x * y - w * z); v02 = T(2) * (x * z + w * y);
		v10 = T(2) * (x * y + w * z); v11 = T(1) - T(2) * (x * x + z * z); v12 = T(2) * (y * z - w * x);
		v20 = T(2) * (x * z - w * y); v21 = T(2) * (y * z + w * x); v22 = T(1) - T(2) * (x * x + y * y);
		// C = (F V0)^T (F V0)
		const T b00 = F[0] * v00 + F[3] * v10 + F[6] * v20, b10 = F[1] * v00 + F[4] * v10 + F[7] * v20, b20 = F[2] * v00 + F[5] * v10 + F[8] * v20;
		const T b01 = F[0] * v01 + F[3] * v11 + F[6] * v21, b11 = F[1] * v01 + F[4] * v11 + F[7] * v21, b21 = F[2] * v01 + F[5] * v11 + F[8] * v21;
		const T b02 = F[0] * v02 + F[3] * v12 + F[6] * v22, b12 = F[1] * v02 + F[4] * v12 + F[7] * v22, b22 = F[2] * v02 + F[5] * v12 + F[8] * v22;
		c00 = b00 * b00 + b10 * b10 + b20 * b20; c01 = b00 * b01 + b10 * b11 + b20 * b21; c02 = b00 * b02 + b10 * b12 + b20 * b22;
		c11 = b01 * b01 + b11 * b11 + b21 * b21; c12 = b01 * b02 + b11 * b12 + b21 * b22; c22 = b02 * b02 + b12 * b12 + b22 * b22;
	} else {
		// C = F^T F
		c00 = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
		c01 = F[0] * F[3] + F[1] * F[4] + F[2] * F[5];
		c02 = F[0] * F[6] + F[1] * F[7] + F[2] * F[8];
		c11 = F[3] * F[3] + F[4] * F[4] + F[5] * F[5];
		c12 = F[3] * F[6] + F[4] * F[7] + F[5] * F[8];
		c22 = F[6] * F[6] + F[7] * F[7] + F[8] * F[8];
	}
#pragma unroll 1
	for (int sweep = 0; sweep < Num<T>::jacobi_sweeps; ++sweep) {
		jacobi_rot(c00, c11, c01, c02, c12, v00, v01, v10, v11, v20, v21); // (0,1), r=2
		jacobi_rot(c00, c22, c02, c01, c12, v00, v02, v10, v12, v20, v22); // (0,2), r=1
		jacobi_rot(c11, c22, c12, c01, c02, v01, v02, v11, v12, v21, v22); // (1,2), r=0
		// converged when every off-diagonal is below the rotation threshold; leave as a warp (all lanes agree)
		const T tr = fabs(c00) + fabs(c11) + fabs(c22);
		const bool done = (fabs(c01) + fabs(c02) + fabs(c12)) <= Num<T>::eps() * T(0.125) * tr;
#if defined(__CUDA_ARCH__)
		if (__all_sync(__activemask(), done)) break;
#else
		if (done) break;
#endif
	}
	// sort eigenvalues descending; a swap of two columns with one negation keeps det V = +1
#define ADMMB200_SWAPCOL(la, lb, a0, a1, a2, b0, b1, b2)                \
	if (la < lb) {                                                        \
		T t_ = la; la = lb; lb = t_;                                      \
		t_ = a0; a0 = b0; b0 = -t_;                                       \
		t_ = a1; a1 = b1; b1 = -t_;                                       \
		t_ = a2; a2 = b2; b2 = -t_;                                       \
	}
	ADMMB200_SWAPCOL(c00, c11, v00, v10, v20, v01, v11, v21)
	ADMMB200_SWAPCOL(c00, c22, v00, v10, v20, v02, v12, v22)
	ADMMB200_SWAPCOL(c11, c22, v01, v11, v21, v02, v12, v22)
#undef ADMMB200_SWAPCOL
	if (q) {
		// V as an (unnormalised) quaternion: of the four equivalent formulas the one with the largest pivot
		const T tw = T(1) + v00 + v11 + v22, tx = T(1) + v00 - v11 - v22, ty = T(1) - v00 + v11 - v22, tz = T(1) - v00 - v11 + v22;
		const T a = v21 - v12, b = v02 - v20, c = v10 - v01, d = v01 + v10, e = v02 + v20, f = v12 + v21;
		const bool pw = tw >= tx && tw >= ty && tw >= tz, px = !pw && tx >= ty && tx >= tz, py = !pw && !px && ty >= tz;
		q[0] = pw ? tw : (px ? a : (py ? b : c));
		q[1] = pw ? a : (px ? tx : (py ? d : e));
		q[2] = pw ? b : (px ? d : (py ? ty : f));
		q[3] = pw ? c : (px ? e : (py ? f : tz));
	}
	// B = F V
	T b00 = F[0] * v00 + F[3] * v10 + F[6] * v20, b10 = F[1] * v00 + F[4] * v10 + F[7] * v20, b20 = F[2] * v00 + F[5] * v10 + F[8] * v20;
	T b01 = F[0] * v01 + F[3] * v11 + F[6] * v21, b11 = F[1] * v01 + F[4] * v11 + F[7] * v21, b21 = F[2] * v01 + F[5] * v11 + F[8] * v21;
	T b02 = F[0] * v02 + F[3] * v12 + F[6] * v22, b12 = F[1] * v02 + F[4] * v12 + F[7] * v22, b22 = F[2] * v02 + F[5] * v12 + F[8] * v22;
	// U by Gram-Schmidt on the (already nearly orthogonal) columns of B, third column = cross product
	// (norm and its reciprocal from ONE reciprocal square root: s = n2 * rsqrt(n2), <= 2 ulp)
	const T n0 = b00 * b00 + b10 * b10 + b20 * b20;
	T s0 = T(0);
	T u00, u10, u20;
	if (n0 > Num<T>::tiny()) { T i = Num<T>::rsqrt(n0); s0 = n0 * i; u00 = b00 * i; u10 = b10 * i; u20 = b20 * i; }
	else { u00 = 1; u10 = 0; u20 = 0; }
	T d = u00 * b01 + u10 * b11 + u20 * b21;
	T w0 = b01 - d * u00, w1 = b11 - d * u10, w2 = b21 - d * u20;
	const T n1 = w0 * w0 + w1 * w1 + w2 * w2;
	T s1 = T(0);
	T u01, u11, u21;
	if (n1 > T(16) * Num<T>::eps() * Num<T>::eps() * n0 && n1 > Num<T>::tiny()) { T i = Num<T>::rsqrt(n1); s1 = n1 * i; u01 = w0 * i; u11 = w1 * i; u21 = w2 * i; }
	else {
		// rank <= 1: any unit vector orthogonal to u0
		T ax = fabs(u00), ay = fabs(u10), az = fabs(u20);
		T e0 = 0, e1 = 0, e2 = 0;
		if (ax <= ay && ax <= az) e0 = 1; else if (ay <= az) e1 = 1; else e2 = 1;
		w0 = u10 * e2 - u20 * e1; w1 = u20 * e0 - u00 * e2; w2 = u00 * e1 - u10 * e0;
		T i = Num<T>::rsqrt(w0 * w0 + w1 * w1 + w2 * w2);
		u01 = w0 * i; u11 = w1 * i; u21 = w2 * i;
		s1 = u01 * b01 + u11 * b11 + u21 * b21;
	}
	T u02 = u10 * u21 - u20 * u11, u12 = u20 * u01 - u00 * u21, u22 = u00 * u11 - u10 * u01;
	T s2 = u02 * b02 + u12 * b12 + u22 * b22;
	S[0] = s0; S[1] = s1; S[2] = s2;
	U[0] = u00; U[1] = u10; U[2] = u20; U[3] = u01; U[4] = u11; U[5] = u21; U[6] = u02; U[7] = u12; U[8] = u22;
	V[0] = v00; V[1] = v10; V[2] = v20; V[3] = v01; V[4] = v11; V[5] = v21; V[6] = v02; V[7] = v12; V[8] = v22;
}

// Z = U diag(s) V^T, column-major.
template <typename T>
ADMMB200_FN void usvt(const T *U, const T *s, const T *V, T *Z)
{
#pragma unroll
	for (int c = 0; c < 3; ++c) {
#pragma unroll
		for (int r = 0; r < 3; ++r) {
			Z[3 * c + r] = U[r] * s[0] * V[c] + U[3 + r] * s[1] * V[3 + c] + U[6 + r] * s[2] * V[6 + c];
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Prox objectives in principal stretches, all divided by K = bulk modulus:
//     phi(x) = Psi(x)/K + 1/2 |x - x0|^2
// (HyperElasticTet::Prox::value = energy_density + k/2 |x-x0|^2; src/TetEnergyTerm.cpp:183-191).
// Each model returns value, gradient g[3] and Hessian h = {h00,h11,h22,h01,h02,h12}.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Material {
	T a, l, kap;              // mu/K, lambda/K, kappa/K (the Newton path works on the objective divided by K)
	double mu, lambda, kappa, k; // raw constants, only read by the reference-faithful path (prox_lbfgs_reference)
	// K = stiffness of the prox penalty K/2 |sigma - sigma0|^2: the bulk modulus of the ELEMENT's Lame
	// (src/TetEnergyTerm.hpp:125-128, 193-200), which for a SplineTet may differ from the spline's own constants;
	// K_ <= 0: lambda + 2/3 mu of the model constants (Lame::bulk_modulus, src/EnergyTerm.hpp:41)
	static Material make(double mu_, double lambda_, double kappa_, double K_ = 0.0) {
		Material m;
		const double K = K_ > 0.0 ? K_ : lambda_ + (2.0 / 3.0) * mu_;
		m.a = T(mu_ / K); m.l = T(lambda_ / K); m.kap = T(kappa_ / K);
		m.mu = mu_; m.lambda = lambda_; m.kappa = kappa_; m.k = K;
		return m;
	}
};

template <typename T, int MODEL> struct Energy;

// NHProx (src/TetEnergyTerm.cpp:173-204): Psi = mu/2 (I1 - log I3 - 3) + lambda/8 log^2 I3
template <typename T> struct Energy<T, TET_NEOHOOKEAN> {
	static ADMMB200_FN T value(const Material<T> &m, const T *x) {
		T lj = Num<T>::log(x[0] * x[1] * x[2]);
		return T(0.5) * m.a * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] - T(2) * lj - T(3)) + T(0.5) * m.l * lj * lj;
	}
	static ADMMB200_FN void derivs(const Material<T> &m, const T *x, T *g, T *h) {
		T lj = Num<T>::log(x[0] * x[1] * x[2]);
		T i0 = Num<T>::frcp(x[0]), i1 = Num<T>::frcp(x[1]), i2 = Num<T>::frcp(x[2]); // <= 1 ulp: 1e-7 relative on g
		T q = m.l * lj - m.a; // (lambda log J - mu)
		g[0] = m.a * x[0] + q * i0; g[1] = m.a * x[1] + q * i1; g[2] = m.a * x[2] + q * i2;
		T p = m.a + m.l - m.l * lj; // mu + lambda (1 - log J)
		h[0] = m.a + p * i0 * i0; h[1] = m.a + p * i1 * i1; h[2] = m.a + p * i2 * i2;
		h[3] = m.l * i0 * i1; h[4] = m.l * i0 * i2; h[5] = m.l * i1 * i2;
	}
};

// StVKProx (src/TetEnergyTerm.cpp:210-237): E = (x^2-1)/2, Psi = mu |E|^2 + lambda/2 tr(E)^2
template <typename T> struct Energy<T, TET_STVK> {
	static ADMMB200_FN T value(const Material<T> &m, const T *x) {
		T e0 = T(0.5) * (x[0] * x[0] - T(1)), e1 = T(0.5) * (x[1] * x[1] - T(1)), e2 = T(0.5) * (x[2] * x[2] - T(1));
		T tr = e0 + e1 + e2;
		return m.a * (e0 * e0 + e1 * e1 + e2 * e2) + T(0.5) * m.l * tr * tr;
	}
	static ADMMB200_FN void derivs(const Material<T> &m, const T *x, T *g, T *h) {
		T n2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
		T q = T(0.5) * m.l * (n2 - T(3));
#pragma unroll
		for (int i = 0; i < 3; ++i) {
			g[i] = m.a * x[i] * (x[i] * x[i] - T(1)) + q * x[i];
			h[i] = m.a * (T(3) * x[i] * x[i] - T(1)) + q + m.l * x[i] * x[i];
		}
		h[3] = m.l * x[0] * x[1]; h[4] = m.l * x[0] * x[2]; h[5] = m.l * x[1] * x[2];
	}
};

// xu::Spline materials (src/XuSpline.hpp:34-94): Psi = sum f(x_i) + sum g(x_i x_j) + h(x0 x1 x2)
// (SplineProx::energy_density, src/TetEnergyTerm.cpp:243-247).  fgh[k] = {value, d, dd}.
template <typename T, int MODEL> struct Spline;
template <typename T> ADMMB200_FN void compress_term(T kap, T x, T *o) {
	T s = (T(1) - x) / T(6);
	o[0] += (kap / T(12)) * s * s * s; o[1] += (-kap / T(24)) * s * s; o[2] += (kap / T(72)) * s;
}
template <typename T> struct Spline<T, TET_SPLINE_NH> {
	static ADMMB200_FN void f(const Material<T> &m, T x, T *o) { o[0] = T(0.5) * m.a * (x * x - T(1)); o[1] = m.a * x; o[2] = m.a; }
	static ADMMB200_FN void g(const Material<T> &, T, T *o) { o[0] = 0; o[1] = 0; o[2] = 0; }
	static ADMMB200_FN void h(const Material<T> &m, T x, T *o) {
		T lx = Num<T>::log(x), ix = T(1) / x;
		o[0] = -m.a * lx + T(0.5) * m.l * lx * lx; o[1] = (m.l * lx - m.a) * ix; o[2] = (m.a + m.l - m.l * lx) * ix * ix;
		compress_term(m.kap, x, o);
	}
};
template <typename T> struct Spline<T, TET_SPLINE_STVK> {
	static ADMMB200_FN void f(const Material<T> &m, T x, T *o) {
		T x2 = x * x;
		o[0] = T(0.125) * m.l * (x2 * x2 - T(6) * x2 + T(5)) + T(0.25) * m.a * (x2 - T(1)) * (x2 - T(1));
		o[1] = T(0.125) * m.l * (T(4) * x2 * x - T(12) * x) + m.a * x * (x2 - T(1));
		o[2] = T(0.125) * m.l * (T(12) * x2 - T(12)) + m.a * (T(3) * x2 - T(1));
	}
	static ADMMB200_FN void g(const Material<T> &m, T x, T *o) { o[0] = T(0.25) * m.l * (x * x - T(1)); o[1] = T(0.5) * m.l * x; o[2] = T(0.5) * m.l; }
	static ADMMB200_FN void h(const Material<T> &m, T x, T *o) { o[0] = 0; o[1] = 0; o[2] = 0; compress_term(m.kap, x, o); }
};
template <typename T> struct Spline<T, TET_SPLINE_COROT> {
	static ADMMB200_FN void f(const Material<T> &m, T x, T *o) {
		o[0] = T(0.5) * m.l * (x * x - T(6) * x + T(5)) + m.a * (x - T(1)) * (x - T(1));
		o[1] = T(0.5) * m.l * (T(2) * x - T(6)) + T(2) * m.a * (x - T(1));
		o[2] = m.l + T(2) * m.a;
	}
	static ADMMB200_FN void g(const Material<T> &m, T x, T *o) { o[0] = m.l * (x - T(1)); o[1] = m.l; o[2] = 0; }
	static ADMMB200_FN void h(const Material<T> &m, T x, T *o) { o[0] = 0; o[1] = 0; o[2] = 0; compress_term(m.kap, x, o); }
};

template <typename T, int MODEL> struct SplineEnergy {
	typedef Spline<T, MODEL> Sp;
	static ADMMB200_FN T value(const Material<T> &m, const T *x) {
		T o[3], v = 0;
		Sp::f(m, x[0], o); v += o[0]; Sp::f(m, x[1], o); v += o[0]; Sp::f(m, x[2], o); v += o[0];
		Sp::g(m, x[0] * x[1], o); v += o[0]; Sp::g(m, x[1] * x[2], o); v += o[0]; Sp::g(m, x[2] * x[0], o); v += o[0];
		Sp::h(m, x[0] * x[1] * x[2], o); v += o[0];
		return v;
	}
	static ADMMB200_FN void derivs(const Material<T> &m, const T *x, T *g, T *h) {
		T f0[3], f1[3], f2[3], g01[3], g12[3], g20[3], hh[3];
		Sp::f(m, x[0], f0); Sp::f(m, x[1], f1); Sp::f(m, x[2], f2);
		Sp::g(m, x[0] * x[1], g01); Sp::g(m, x[1] * x[2], g12); Sp::g(m, x[2] * x[0], g20);
		Sp::h(m, x[0] * x[1] * x[2], hh);
		T y0 = x[1] * x[2], y1 = x[2] * x[0], y2 = x[0] * x[1]; // dJ/dx_i
		g[0] = f0[1] + g01[1] * x[1] + g20[1] * x[2] + hh[1] * y0;
		g[1] = f1[1] + g12[1] * x[2] + g01[1] * x[0] + hh[1] * y1;
		g[2] = f2[1] + g20[1] * x[0] + g12[1] * x[1] + hh[1] * y2;
		h[0] = f0[2] + g01[2] * x[1] * x[1] + g20[2] * x[2] * x[2] + hh[2] * y0 * y0;
		h[1] = f1[2] + g12[2] * x[2] * x[2] + g01[2] * x[0] * x[0] + hh[2] * y1 * y1;
		h[2] = f2[2] + g20[2] * x[0] * x[0] + g12[2] * x[1] * x[1] + hh[2] * y2 * y2;
		h[3] = g01[2] * y2 + g01[1] + hh[2] * y0 * y1 + hh[1] * x[2];
		h[4] = g20[2] * y1 + g20[1] + hh[2] * y0 * y2 + hh[1] * x[1];
		h[5] = g12[2] * y0 + g12[1] + hh[2] * y1 * y2 + hh[1] * x[0];
	}
};
template <typename T> struct Energy<T, TET_SPLINE_NH> : SplineEnergy<T, TET_SPLINE_NH> {};
template <typename T> struct Energy<T, TET_SPLINE_STVK> : SplineEnergy<T, TET_SPLINE_STVK> {};
template <typename T> struct Energy<T, TET_SPLINE_COROT> : SplineEnergy<T, TET_SPLINE_COROT> {};

// does the model need x > 0 strictly (log barrier)?
template <int MODEL> struct NeedsPositive { static constexpr bool value = (MODEL == TET_NEOHOOKEAN || MODEL == TET_SPLINE_NH); };

// argmin_x>=0 phi(x), started at x (already made feasible), quadratic centre x0.
// Returns true when the result is NOT a converged interior minimiser (a stretch ended on the bound
// x = 0, or the iteration budget ran out): the caller then reproduces the reference's own optimiser.
template <typename T, int MODEL>
ADMMB200_FN bool prox_newton(const Material<T> &m, const T *x0, T *x)
{
	bool converged = false;
	typedef Energy<T, MODEL> En;
	const T floor_x = NeedsPositive<MODEL>::value ? T(1e-12) : T(0);
	T phi_x = T(0);          // phi at x, carried from the accepted trial point of the previous iteration
	bool have_phi = false;
#pragma unroll 1
	for (int it = 0; it < Num<T>::newton_iters; ++it) {
		T g[3], h[6];
		En::derivs(m, x, g, h);
		g[0] += x[0] - x0[0]; g[1] += x[1] - x0[1]; g[2] += x[2] - x0[2];
		h[0] += T(1); h[1] += T(1); h[2] += T(1);
		// Active set of the bound x >= 0 (value() is +inf for x<0, src/TetEnergyTerm.cpp:184-188): a
		// stretch that sits on the bound with the objective still decreasing outwards stays there and
		// the Newton system is solved for the free stretches only (inverted StVK / spline elements end
		// on the face x2 = 0, where the reference's barrier line search creeps to as well).
		if (!NeedsPositive<MODEL>::value) {
			const T on_bound = T(1e-7);
			if (x[0] <= on_bound && g[0] > T(0)) { x[0] = T(0); g[0] = T(0); h[0] = T(1); h[3] = T(0); h[4] = T(0); }
			if (x[1] <= on_bound && g[1] > T(0)) { x[1] = T(0); g[1] = T(0); h[1] = T(1); h[3] = T(0); h[5] = T(0); }
			if (x[2] <= on_bound && g[2] > T(0)) { x[2] = T(0); g[2] = T(0); h[2] = T(1); h[4] = T(0); h[5] = T(0); }
		}
		// Newton direction by LDL^T (reciprocals of the pivots are approximate: they only steer the
		// iteration); if H is not positive definite fall back to a scaled gradient step
		T d[3];
		const T d0 = h[0];
		const T r0 = Num<T>::frcp(d0);
		const T l10 = h[3] * r0, l20 = h[4] * r0;
		const T d1 = h[1] - l10 * h[3];
		const T r1 = Num<T>::frcp(d1);
		const T l21 = (h[5] - l20 * h[3]) * r1;
		const T d2 = h[2] - l20 * h[4] - l21 * l21 * d1;
		const bool pd = (d0 > T(0)) && (d1 > T(0)) && (d2 > T(0));
		if (pd) {
			const T r2 = Num<T>::frcp(d2);
			T y0 = -g[0], y1 = -g[1] - l10 * y0, y2 = -g[2] - l20 * y0 - l21 * y1;
			T z2 = y2 * r2, z1 = y1 * r1 - l21 * z2, z0 = y0 * r0 - l10 * z1 - l20 * z2;
			d[0] = z0; d[1] = z1; d[2] = z2;
		} else {
			T bound = fmax(fmax(fabs(h[0]) + fabs(h[3]) + fabs(h[4]), fabs(h[1]) + fabs(h[3]) + fabs(h[5])), fabs(h[2]) + fabs(h[4]) + fabs(h[5]));
			T sc = T(1) / fmax(bound, T(1));
			d[0] = -g[0] * sc; d[1] = -g[1] * sc; d[2] = -g[2] * sc;
		}
		// keep the iterate inside the feasible set
		T t = T(1);
#pragma unroll
		for (int i = 0; i < 3; ++i) {
			if (x[i] + d[i] < floor_x) {
				T room = (x[i] - floor_x);
				T ti = (NeedsPositive<MODEL>::value ? T(0.9) : T(1)) * room / (-d[i]);
				t = fmin(t, ti);
			}
		}
		const T dmax = fmax(fmax(fabs(d[0]), fabs(d[1])), fabs(d[2]));
		const T xmin = fmin(fmin(x[0], x[1]), x[2]);
		T xn[3];
		// A full Newton step that is small against the distance to the bound and taken with a positive
		// definite Hessian needs no safeguard; everything else goes through the Armijo backtracking.
		if (pd && t == T(1) && dmax < T(0.05) * xmin) {
			xn[0] = x[0] + d[0]; xn[1] = x[1] + d[1]; xn[2] = x[2] + d[2];
			have_phi = false;
		} else {
			T gd = g[0] * d[0] + g[1] * d[1] + g[2] * d[2];
			if (!have_phi) {
				T dx0 = x[0] - x0[0], dx1 = x[1] - x0[1], dx2 = x[2] - x0[2];
				phi_x = En::value(m, x) + T(0.5) * (dx0 * dx0 + dx1 * dx1 + dx2 * dx2);
			}
			T slack = T(8) * Num<T>::eps() * (fabs(phi_x) + T(1));
			T phi = phi_x;
#pragma unroll 1
			for (int ls = 0; ls < 16; ++ls) {
				xn[0] = fmax(x[0] + t * d[0], floor_x); xn[1] = fmax(x[1] + t * d[1], floor_x); xn[2] = fmax(x[2] + t * d[2], floor_x);
				T e0 = xn[0] - x0[0], e1 = xn[1] - x0[1], e2 = xn[2] - x0[2];
				phi = En::value(m, xn) + T(0.5) * (e0 * e0 + e1 * e1 + e2 * e2);
				if (phi <= phi_x + T(1e-4) * t * gd + slack) break;
				t *= T(0.5);
			}
			phi_x = phi; have_phi = true;
		}
		const T step = t * dmax;
		x[0] = xn[0]; x[1] = xn[1]; x[2] = xn[2];
		const T scale = T(1) + fmax(fmax(fabs(x[0]), fabs(x[1])), fabs(x[2]));
		// quadratic convergence: after a (safeguard-free or full) step of size s the error is O(s^2)
		if (step <= Num<T>::newton_tol() * scale && t == T(1) && pd) { converged = true; break; }
		if (step <= T(4) * Num<T>::eps() * scale) { converged = true; break; }
	}
	// "near the bound" is generous on purpose: the slow path IS the reference's algorithm
	return !converged || (!NeedsPositive<MODEL>::value && (x[0] < T(1e-3) || x[1] < T(1e-3) || x[2] < T(1e-3)));
}

// ---------------------------------------------------------------------------------------------
// Reference-faithful path for DEGENERATE elements only.
//
// When the minimiser of the prox objective lies on the bound x = 0 (inverted or nearly flat StVK /
// spline elements) the reference does not reach it: its L-BFGS creeps towards the bound through a
// line search whose trial points beyond x = 0 evaluate to FLT_MAX, and stops by |dx| < 1e-6
// (src/TetEnergyTerm.hpp:93-95).  The answer is then a property of the optimiser's path, so the only
// way to agree with it is to walk the same path: LBFGS<double,3,8>::minimize
// (deps/mcloptlib/include/MCL/LBFGS.hpp:52-152: history 8, max 50 iterations, steepest-descent restart
// with alpha = min(1, 1/|g|_inf) on a non-descent direction, FAILURE on rate <= 0) with
// BacktrackingCubic::search (Backtracking.hpp:79-143; ls_decrease 1e-4, ls_max_iters 100000,
// Minimizer.hpp:66-70), in fp64 and in the reference's unscaled units.  Out of line and never inlined:
// the hot path pays one predicted-not-taken branch.
// ---------------------------------------------------------------------------------------------
template <int MODEL> struct RefProblem {
	Material<double> m; // a = mu, l = lambda, kap = kappa (raw)
	double k, x0[3];
	ADMMB200_FN double value(const double *x) const { // Prox::value (src/TetEnergyTerm.cpp:184-192, 210-218, 249-257)
		if (x[0] < 0.0 || x[1] < 0.0 || x[2] < 0.0) return 3.4028234663852886e38; // numeric_limits<float>::max()
		double d0 = x[0] - x0[0], d1 = x[1] - x0[1], d2 = x[2] - x0[2];
		return Energy<double, MODEL>::value(m, x) + (k * 0.5) * (d0 * d0 + d1 * d1 + d2 * d2);
	}
	ADMMB200_FN double gradient(const double *x, double *g) const { // Prox::gradient (:195-204, 228-237, 259-265)
		double h[6];
		Energy<double, MODEL>::derivs(m, x, g, h);
		g[0] += k * (x[0] - x0[0]); g[1] += k * (x[1] - x0[1]); g[2] += k * (x[2] - x0[2]);
		return value(x);
	}
};

template <int MODEL>
ADMMB200_SLOWPATH void prox_lbfgs_reference(double mu, double lambda, double kappa, double bulk, const double *x0, double *x)
{
	RefProblem<MODEL> P;
	P.m.a = mu; P.m.l = lambda; P.m.kap = kappa; P.m.mu = mu; P.m.lambda = lambda; P.m.kappa = kappa;
	P.k = bulk;
	P.x0[0] = x0[0]; P.x0[1] = x0[1]; P.x0[2] = x0[2];
	const int M = 8;
	double s[M][3], y[M][3], alpha[M], rho[M];
	double grad[3], q[3], grad_old[3], x_old[3];
	double gamma_k = 1.0, alpha_init = 1.0;
	int max_iters = 50;
	for (int i = 0; i < M; ++i) for (int j = 0; j < 3; ++j) { s[i][j] = 0.0; y[i][j] = 0.0; }
	if (NeedsPositive<MODEL>::value && x[0] * x[1] * x[2] <= 0.0) return; // the reference's gradient() throws here
	P.gradient(x, grad);
	for (int k = 0; k < max_iters; ++k) {
		const int iter = k < M ? k : M;
		for (int j = 0; j < 3; ++j) { x_old[j] = x[j]; grad_old[j] = grad[j]; q[j] = grad[j]; }
		for (int i = iter - 1; i >= 0; --i) { // two-loop recursion (LBFGS.hpp:87-100)
			rho[i] = 1.0 / (s[i][0] * y[i][0] + s[i][1] * y[i][1] + s[i][2] * y[i][2]);
			alpha[i] = rho[i] * (s[i][0] * q[0] + s[i][1] * q[1] + s[i][2] * q[2]);
			for (int j = 0; j < 3; ++j) q[j] -= alpha[i] * y[i][j];
		}
		for (int j = 0; j < 3; ++j) q[j] *= gamma_k;
		for (int i = 0; i < iter; ++i) {
			double beta = rho[i] * (q[0] * y[i][0] + q[1] * y[i][1] + q[2] * y[i][2]);
			for (int j = 0; j < 3; ++j) q[j] += (alpha[i] - beta) * s[i][j];
		}
		if (q[0] * grad[0] + q[1] * grad[1] + q[2] * grad[2] <= 0.0) { // not a descent direction: restart (:102-109)
			double inf = fmax(fabs(grad[0]), fmax(fabs(grad[1]), fabs(grad[2])));
			for (int j = 0; j < 3; ++j) q[j] = grad[j];
			max_iters -= k; k = 0;
			alpha_init = fmin(1.0, 1.0 / inf);
		}
		// BacktrackingCubic::search along -q
		double dir[3] = {-q[0], -q[1], -q[2]};
		double rate;
		if (sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]) <= 2.220446049250313e-16) rate = 1e-4;
		else {
			double g0[3];
			if (NeedsPositive<MODEL>::value && x[0] * x[1] * x[2] <= 0.0) return;
			const double fx0 = P.gradient(x, g0);
			const double gtp = g0[0] * dir[0] + g0[1] * dir[1] + g0[2] * dir[2];
			double fxp = fx0, a = alpha_init, ap = alpha_init;
			int it = 0;
			const int ls_max = 100000;
			for (; it < ls_max; ++it) {
				double xa[3] = {x[0] + a * dir[0], x[1] + a * dir[1], x[2] + a * dir[2]};
				double fxa = P.value(xa);
				if (fxa <= fx0 + a * 1e-4 * gtp) break;
				double at;
				if (it == 0) at = gtp / (2.0 * (fx0 + gtp - fxa));
				else { // cubic interpolation through the last two trials (Backtracking.hpp:129-143)
					double mult = 1.0 / (a * a * ap * ap * (a - ap));
					double B0 = fxa - fx0 - a * gtp, B1 = fxp - fx0 - ap * gtp;
					double r0 = mult * (ap * ap * B0 - a * a * B1), r1 = mult * (-ap * ap * ap * B0 + a * a * a * B1);
					if (fabs(r0) <= 0.0) at = -gtp / (2.0 * r1);
					else at = (-r1 + sqrt(r1 * r1 - 3.0 * r0 * gtp)) / (3.0 * r0);
				}
				fxp = fxa; ap = a;
				if (at != at) a = at; // range() hands a NaN through
				else { double lo = 0.1 * a, hi = 0.5 * a; a = at < lo ? lo : (at > hi ? hi : at); }
			}
			rate = it >= ls_max ? -1.0 : a;
		}
		if (!(rate > 0.0)) return; // Minimizer::FAILURE: x stays at the last iterate (:113-116)
		double xl[3] = {x[0], x[1], x[2]};
		for (int j = 0; j < 3; ++j) x[j] -= rate * q[j];
		{ // converged(x_last, x, grad of the previous iterate) (:120, src/TetEnergyTerm.hpp:93-95)
			double d0 = xl[0] - x[0], d1 = xl[1] - x[1], d2 = xl[2] - x[2];
			if (sqrt(grad[0] * grad[0] + grad[1] * grad[1] + grad[2] * grad[2]) < 1e-6 || sqrt(d0 * d0 + d1 * d1 + d2 * d2) < 1e-6) break;
		}
		if (NeedsPositive<MODEL>::value && x[0] * x[1] * x[2] <= 0.0) return;
		P.gradient(x, grad);
		double st[3], yt[3];
		for (int j = 0; j < 3; ++j) { st[j] = x[j] - x_old[j]; yt[j] = grad[j] - grad_old[j]; }
		if (k < M) { for (int j = 0; j < 3; ++j) { s[k][j] = st[j]; y[k][j] = yt[j]; } }
		else {
			for (int i = 0; i < M - 1; ++i) for (int j = 0; j < 3; ++j) { s[i][j] = s[i + 1][j]; y[i][j] = y[i + 1][j]; }
			for (int j = 0; j < 3; ++j) { s[M - 1][j] = st[j]; y[M - 1][j] = yt[j]; }
		}
		double denom = yt[0] * yt[0] + yt[1] * yt[1] + yt[2] * yt[2];
		if (fabs(denom) <= 0.0) break;
		gamma_k = (st[0] * yt[0] + st[1] * yt[1] + st[2] * yt[2]) / denom;
		alpha_init = 1.0;
	}
}

// HyperElasticTet::prox (src/TetEnergyTerm.cpp:114-136) on a column-major 3x3 z (in/out).
//   PROX_INLINE     Newton, and the reference-faithful path right here when the element is degenerate
//   PROX_FAST       Newton only; returns true (z unspecified) for a degenerate element, so that the
//                   caller can queue it -- the hot kernel then carries no call, no stack, fewer registers
//   PROX_REFERENCE  the reference-faithful path only (the queue's consumer)
enum ProxMode { PROX_INLINE = 0, PROX_FAST = 1, PROX_REFERENCE = 2 };

template <typename T, int MODEL, int MODE>
ADMMB200_FN bool prox_tet_mode(const Material<T> &m, T *z, T *q = nullptr)
{
	T S[3], U[9], V[9];
	svd3_signed(z, S, U, V, q);
	if (MODEL == TET_LINEAR) {
		// TetEnergyTerm::prox (src/TetEnergyTerm.cpp:73-92): p = U diag(1,1,sign det F) V^T with
		// Eigen's unsigned factors = U V^T with the proper rotations computed here; z = (p+z)/2.
		T one[3] = {T(1), T(1), T(1)}, P[9];
		usvt(U, one, V, P);
#pragma unroll
		for (int i = 0; i < 9; ++i) z[i] = T(0.5) * (P[i] + z[i]);
		return false;
	}
	T x0[3] = {S[0], S[1], S[2]};
	const T eps = T(1e-6);
	bool collapsed = false;
	if (fabs(S[0]) < eps && fabs(S[1]) < eps && fabs(S[2]) < eps) { S[0] = eps; S[1] = eps; S[2] = eps; collapsed = true; }
	if (S[2] < T(0)) S[2] = -S[2];
	if (NeedsPositive<MODEL>::value) {
		// the start point of the reference can sit exactly on the barrier (S[2]==0); nudge it inside
		const T lo = T(1e-7);
		S[0] = fmax(S[0], lo); S[1] = fmax(S[1], lo); S[2] = fmax(S[2], lo);
	}
	// an element collapsed to a point starts 6 decades from its minimiser: where the reference's
	// optimiser stops from there is path-dependent too (src/TetEnergyTerm.cpp:126-129)
	bool degenerate = collapsed;
	if (MODE == PROX_REFERENCE) degenerate = true;
	else if (!collapsed) {
		T start[3] = {S[0], S[1], S[2]};
		if (MODEL == TET_NEOHOOKEAN) degenerate = prox_newton<T, TET_NEOHOOKEAN>(m, x0, S);
		if (MODEL == TET_STVK) degenerate = prox_newton<T, TET_STVK>(m, x0, S);
		if (MODEL == TET_SPLINE_NH) degenerate = prox_newton<T, TET_SPLINE_NH>(m, x0, S);
		if (MODEL == TET_SPLINE_STVK) degenerate = prox_newton<T, TET_SPLINE_STVK>(m, x0, S);
		if (MODEL == TET_SPLINE_COROT) degenerate = prox_newton<T, TET_SPLINE_COROT>(m, x0, S);
		if (degenerate) { S[0] = start[0]; S[1] = start[1]; S[2] = start[2]; }
	}
	if (degenerate) {
		if (MODE == PROX_FAST) return true;
		// minimiser on the bound (or no convergence): walk the reference optimiser's own path in fp64
		double xs[3] = {double(S[0]), double(S[1]), double(S[2])}, xc[3] = {double(x0[0]), double(x0[1]), double(x0[2])};
		if (MODEL != TET_LINEAR) prox_lbfgs_reference<MODEL == TET_LINEAR ? TET_STVK : MODEL>(m.mu, m.lambda, m.kappa, m.k, xc, xs);
		S[0] = T(xs[0]); S[1] = T(xs[1]); S[2] = T(xs[2]);
	}
	usvt(U, S, V, z);
	return false;
}

template <typename T, int MODEL>
ADMMB200_FN void prox_tet(const Material<T> &m, T *z) { prox_tet_mode<T, MODEL, PROX_INLINE>(m, z); }

// TriEnergyTerm::prox (src/TriEnergyTerm.cpp:73-101) on a column-major 3x2 z (in/out):
// P = U[:, :2] V^T (nearest matrix with singular values 1,1), z = (P+z)/2, then the optional
// column-norm strain limit.
template <typename T>
ADMMB200_FN void prox_tri(T limit_min, T limit_max, T *z)
{
	T a = z[0] * z[0] + z[1] * z[1] + z[2] * z[2];
	T b = z[0] * z[3] + z[1] * z[4] + z[2] * z[5];
	T d = z[3] * z[3] + z[4] * z[4] + z[5] * z[5];
	T c = 1, s = 0;
	if (fabs(b) > Num<T>::eps() * T(0.125) * (a + d) && fabs(b) > Num<T>::tiny()) {
		T theta = (d - a) / (T(2) * b);
		T t = copysign(T(1), theta) / (fabs(theta) + Num<T>::sqrt(theta * theta + T(1)));
		c = Num<T>::rsqrt(t * t + T(1)); s = t * c;
	}
	// V = [c s; -s c] (columns v0 = (c,-s), v1 = (s,c)); B = F V
	T b0[3] = {c * z[0] - s * z[3], c * z[1] - s * z[4], c * z[2] - s * z[5]};
	T b1[3] = {s * z[0] + c * z[3], s * z[1] + c * z[4], s * z[2] + c * z[5]};
	T n0 = b0[0] * b0[0] + b0[1] * b0[1] + b0[2] * b0[2];
	T n1 = b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2];
	if (n0 < n1) { // order so that u0 belongs to the larger singular value (better conditioned first)
		T t;
		for (int i = 0; i < 3; ++i) { t = b0[i]; b0[i] = b1[i]; b1[i] = -t; }
		t = c; T s_old = s; c = s_old; s = -t; // new v0 = old v1 = (s,c), new v1 = -old v0 = (-c, s)
		t = n0; n0 = n1; n1 = t;
	}
	T u0[3], u1[3];
	T s0 = Num<T>::sqrt(n0);
	if (s0 > Num<T>::tiny()) { T i = T(1) / s0; u0[0] = b0[0] * i; u0[1] = b0[1] * i; u0[2] = b0[2] * i; }
	else { u0[0] = 1; u0[1] = 0; u0[2] = 0; }
	T dd = u0[0] * b1[0] + u0[1] * b1[1] + u0[2] * b1[2];
	T w[3] = {b1[0] - dd * u0[0], b1[1] - dd * u0[1], b1[2] - dd * u0[2]};
	T s1 = Num<T>::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
	if (s1 > T(4) * Num<T>::eps() * s0 && s1 > Num<T>::tiny()) { T i = T(1) / s1; u1[0] = w[0] * i; u1[1] = w[1] * i; u1[2] = w[2] * i; }
	else {
		T ax = fabs(u0[0]), ay = fabs(u0[1]), az = fabs(u0[2]);
		T e[3] = {0, 0, 0};
		if (ax <= ay && ax <= az) e[0] = 1; else if (ay <= az) e[1] = 1; else e[2] = 1;
		w[0] = u0[1] * e[2] - u0[2] * e[1]; w[1] = u0[2] * e[0] - u0[0] * e[2]; w[2] = u0[0] * e[1] - u0[1] * e[0];
		T i = Num<T>::rsqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
		u1[0] = w[0] * i; u1[1] = w[1] * i; u1[2] = w[2] * i;
	}
	// P = u0 v0^T + u1 v1^T with v0 = (c,-s), v1 = (s,c);  column 0 of P = u0*c + u1*s, column 1 = -u0*s + u1*c
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		T p0 = u0[i] * c + u1[i] * s;
		T p1 = -u0[i] * s + u1[i] * c;
		z[i] = T(0.5) * (p0 + z[i]);
		z[3 + i] = T(0.5) * (p1 + z[3 + i]);
	}
	const bool check_strain = limit_min > T(0) || limit_max < T(99);
	if (check_strain) {
		T l0 = Num<T>::sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
		T l1 = Num<T>::sqrt(z[3] * z[3] + z[4] * z[4] + z[5] * z[5]);
		if (l0 < limit_min) { T f = limit_min / l0; z[0] *= f; z[1] *= f; z[2] *= f; }
		if (l1 < limit_min) { T f = limit_min / l1; z[3] *= f; z[4] *= f; z[5] *= f; }
		if (l0 > limit_max) { T f = limit_max / l0; z[0] *= f; z[1] *= f; z[2] *= f; }
		if (l1 > limit_max) { T f = limit_max / l1; z[3] *= f; z[4] *= f; z[5] *= f; }
	}
}

} // namespace admmb200
