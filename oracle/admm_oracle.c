/* oracle/admm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, fp64, CPU restatement of the hot path of mattoverby/admm-elastic @ c6c09a3: one
 * admm::Solver::step() with its local step (EnergyTerm::update + the prox of every term type), the
 * right-hand-side assembly and the three global solvers.  Every function cites the reference
 * file:line it follows.  It exists to CHECK the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product never calls it.
 *
 * Pinning (tests/test_oracle_*.py, run without a GPU):
 *   - the reference's own known answers in samples/tests/test_lineartet.cpp (F layout, energies,
 *     x = 52.2321 +- 1e-4, inversion recovery), and
 *   - outputs of the reference itself, compiled unmodified into oracle/_ref/libadmm_ref.so, on
 *     seeded inputs (prox vectors of every model, full steps with LDLT / MCGS / Uzawa), and the
 *     golden fixtures generated from it under tests/golden/.
 *
 * Where this file deliberately differs from the reference:
 *   - the 3x3 / 3x2 SVD is a one-sided Jacobi written here, not Eigen::JacobiSVD; the products
 *     U f(S) V^T agree to rounding (the reference's FastSVD.hpp is itself a stub around Eigen);
 *   - the colouring of NodalMultiColorGS is an INPUT (oracle_set_colors): the reference's is
 *     randomised and time-seeded (GraphColor.hpp:156), SURVEY.md 0.5;
 *   - the sparse LDL^T uses the natural ordering instead of Eigen's AMD: same solution to rounding;
 *   - exceptions become error codes; SpringPin's out-of-bounds read is not replicated.
 */
#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------- */
/* small dense helpers                                                                          */
/* ------------------------------------------------------------------------------------------- */
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double *a) { return sqrt(dot3(a, a)); }
static void cross3(const double *a, const double *b, double *c) { c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; }
static double det3cm(const double *F) { /* column-major */
	return F[0] * (F[4] * F[8] - F[7] * F[5]) - F[3] * (F[1] * F[8] - F[7] * F[2]) + F[6] * (F[1] * F[5] - F[4] * F[2]);
}

/* One-sided (Hestenes) Jacobi SVD of an m x n column-major matrix, n <= 3: A = U diag(S) V^T with
 * S >= 0 sorted descending.  Stands in for Eigen::JacobiSVD (src/FastSVD.hpp:47, src/TetEnergyTerm.cpp:76,
 * src/TriEnergyTerm.cpp:78).  U is m x n (thin). */
static void jacobi_svd(int m, int n, const double *A, double *U, double *S, double *V)
{
	double B[9];
	int i, j, k, sweep;
	memcpy(B, A, sizeof(double) * m * n);
	for (i = 0; i < n * n; ++i) V[i] = 0.0;
	for (i = 0; i < n; ++i) V[i * n + i] = 1.0;
	for (sweep = 0; sweep < 60; ++sweep) {
		int rotated = 0;
		for (i = 0; i < n - 1; ++i) for (j = i + 1; j < n; ++j) {
			double a = 0, b = 0, c = 0;
			for (k = 0; k < m; ++k) { a += B[i * m + k] * B[i * m + k]; b += B[j * m + k] * B[j * m + k]; c += B[i * m + k] * B[j * m + k]; }
			if (fabs(c) <= 1e-300 || fabs(c) <= DBL_EPSILON * 0.01 * sqrt(a * b)) continue;
			{
				double zeta = (b - a) / (2.0 * c);
				double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
				double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
				for (k = 0; k < m; ++k) { double bi = B[i * m + k], bj = B[j * m + k]; B[i * m + k] = cs * bi - sn * bj; B[j * m + k] = sn * bi + cs * bj; }
				for (k = 0; k < n; ++k) { double vi = V[i * n + k], vj = V[j * n + k]; V[i * n + k] = cs * vi - sn * vj; V[j * n + k] = sn * vi + cs * vj; }
				rotated = 1;
			}
		}
		if (!rotated) break;
	}
	for (i = 0; i < n; ++i) { double s = 0; for (k = 0; k < m; ++k) s += B[i * m + k] * B[i * m + k]; S[i] = sqrt(s); }
	/* sort descending (selection sort on columns) */
	for (i = 0; i < n - 1; ++i) {
		int best = i;
		for (j = i + 1; j < n; ++j) if (S[j] > S[best]) best = j;
		if (best != i) {
			double t = S[i]; S[i] = S[best]; S[best] = t;
			for (k = 0; k < m; ++k) { t = B[i * m + k]; B[i * m + k] = B[best * m + k]; B[best * m + k] = t; }
			for (k = 0; k < n; ++k) { t = V[i * n + k]; V[i * n + k] = V[best * n + k]; V[best * n + k] = t; }
		}
	}
	/* U columns; complete degenerate directions orthogonally */
	for (i = 0; i < n; ++i) {
		if (S[i] > 1e-300 && (i == 0 || S[i] > 4.0 * DBL_EPSILON * S[0])) { for (k = 0; k < m; ++k) U[i * m + k] = B[i * m + k] / S[i]; }
		else {
			/* m == 3 always here */
			double e[3], w[3], best = -1; int pick = 0, tries;
			if (i == 0) { U[0] = 1; U[1] = 0; U[2] = 0; continue; }
			if (i == 1) {
				for (tries = 0; tries < 3; ++tries) { if (best < 0 || fabs(U[tries]) < best) { best = fabs(U[tries]); pick = tries; } }
				e[0] = e[1] = e[2] = 0; e[pick] = 1;
				cross3(U, e, w);
			} else { cross3(U, U + 3, w); }
			{ double nw = norm3(w); for (k = 0; k < 3; ++k) U[i * m + k] = w[k] / nw; }
		}
	}
}

/* signed_svd (src/FastSVD.hpp:43-68): U, V forced into SO(3), the sign goes to S[2]. Column-major. */
static void signed_svd(const double *F, double *S, double *U, double *V)
{
	int k;
	jacobi_svd(3, 3, F, U, S, V);
	if (det3cm(U) < 0.0) { for (k = 0; k < 3; ++k) U[6 + k] = -U[6 + k]; S[2] = -S[2]; }
	if (det3cm(V) < 0.0) { for (k = 0; k < 3; ++k) V[6 + k] = -V[6 + k]; S[2] = -S[2]; }
}

static void usvt(const double *U, const double *s, const double *V, double *Z)
{
	int r, c;
	for (c = 0; c < 3; ++c) for (r = 0; r < 3; ++r) Z[3 * c + r] = U[r] * s[0] * V[c] + U[3 + r] * s[1] * V[3 + c] + U[6 + r] * s[2] * V[6 + c];
}

/* ------------------------------------------------------------------------------------------- */
/* prox objectives (src/TetEnergyTerm.cpp:173-265, src/XuSpline.hpp:34-94)                      */
/* ------------------------------------------------------------------------------------------- */
enum { TET_LINEAR = 0, TET_NEOHOOKEAN = 1, TET_STVK = 2, TET_SPLINE_NH = 3, TET_SPLINE_STVK = 4, TET_SPLINE_COROT = 5 };

typedef struct { int model; double mu, lambda, kappa, k; double x0[3]; int error; } prox_problem;

/* xu::Spline::compress_term / d_compress_term (src/XuSpline.hpp:43-45) */
static double compress_term(double kappa, double x) { return (kappa / 12.0) * pow((1.0 - x) / 6.0, 3.0); }
static double d_compress_term(double kappa, double x) { return (-kappa / 24.0) * pow((1.0 - x) / 6.0, 2.0); }

static double sp_f(const prox_problem *p, double x) {
	double x2 = x * x;
	switch (p->model) {
	case TET_SPLINE_NH: return 0.5 * p->mu * (x * x - 1.0);                                                      /* XuSpline.hpp:53 */
	case TET_SPLINE_STVK: return 0.125 * p->lambda * (x2 * x2 - 6.0 * x2 + 5.0) + 0.25 * p->mu * (x2 - 1.0) * (x2 - 1.0); /* :69-72 */
	default: return 0.5 * p->lambda * (x * x - 6.0 * x + 5.0) + p->mu * (x - 1.0) * (x - 1.0);                   /* :88 */
	}
}
static double sp_g(const prox_problem *p, double x) {
	switch (p->model) {
	case TET_SPLINE_NH: return 0.0;
	case TET_SPLINE_STVK: return 0.25 * p->lambda * (x * x - 1.0);
	default: return p->lambda * (x - 1.0);
	}
}
static double sp_h(const prox_problem *p, double x) {
	if (p->model == TET_SPLINE_NH) { double lx = log(x); return -p->mu * lx + 0.5 * p->lambda * lx * lx + compress_term(p->kappa, x); }
	return compress_term(p->kappa, x);
}
static double sp_df(const prox_problem *p, double x) {
	double x2 = x * x;
	switch (p->model) {
	case TET_SPLINE_NH: return p->mu * x;
	case TET_SPLINE_STVK: return 0.125 * p->lambda * (4.0 * x2 * x - 12.0 * x) + p->mu * x * (x2 - 1.0);
	default: return 0.5 * p->lambda * (2.0 * x - 6.0) + 2.0 * p->mu * (x - 1.0);
	}
}
static double sp_dg(const prox_problem *p, double x) {
	switch (p->model) {
	case TET_SPLINE_NH: return 0.0;
	case TET_SPLINE_STVK: return 0.5 * p->lambda * x;
	default: return p->lambda;
	}
}
static double sp_dh(const prox_problem *p, double x) {
	if (p->model == TET_SPLINE_NH) return -p->mu / x + p->lambda * log(x) / x + d_compress_term(p->kappa, x);
	return d_compress_term(p->kappa, x);
}

static double energy_density(const prox_problem *p, const double *x)
{
	switch (p->model) {
	case TET_NEOHOOKEAN: { /* NHProx::energy_density, src/TetEnergyTerm.cpp:173-182 */
		double J = x[0] * x[1] * x[2], I1 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2], I3 = J * J, l = log(I3);
		return 0.5 * p->mu * (I1 - l - 3.0) + 0.125 * p->lambda * l * l;
	}
	case TET_STVK: { /* StVKProx::energy_density, :220-226 */
		double s0 = 0.5 * (x[0] * x[0] - 1.0), s1 = 0.5 * (x[1] * x[1] - 1.0), s2 = 0.5 * (x[2] * x[2] - 1.0), tr = s0 + s1 + s2;
		return p->mu * (s0 * s0 + s1 * s1 + s2 * s2) + p->lambda * 0.5 * tr * tr;
	}
	default: /* SplineProx::energy_density, :243-247 */
		return sp_f(p, x[0]) + sp_f(p, x[1]) + sp_f(p, x[2]) + sp_g(p, x[0] * x[1]) + sp_g(p, x[1] * x[2]) + sp_g(p, x[2] * x[0]) + sp_h(p, x[0] * x[1] * x[2]);
	}
}

/* Prox::value (src/TetEnergyTerm.cpp:184-192, 210-218, 249-257) */
static double prox_value(prox_problem *p, const double *x)
{
	double d0, d1, d2;
	if (x[0] < 0.0 || x[1] < 0.0 || x[2] < 0.0) return (double)FLT_MAX;
	d0 = x[0] - p->x0[0]; d1 = x[1] - p->x0[1]; d2 = x[2] - p->x0[2];
	return energy_density(p, x) + (p->k * 0.5) * (d0 * d0 + d1 * d1 + d2 * d2);
}

/* Prox::gradient (src/TetEnergyTerm.cpp:195-204, 228-237, 259-265) */
static double prox_gradient(prox_problem *p, const double *x, double *g)
{
	int i;
	switch (p->model) {
	case TET_NEOHOOKEAN: {
		double J = x[0] * x[1] * x[2];
		if (J <= 0.0) { p->error = 1; g[0] = g[1] = g[2] = 0; return (double)FLT_MAX; } /* the reference throws here */
		for (i = 0; i < 3; ++i) { double xi = 1.0 / x[i]; g[i] = (p->mu * (x[i] - xi) + p->lambda * log(J) * xi) + p->k * (x[i] - p->x0[i]); }
	} break;
	case TET_STVK: {
		double xx = dot3(x, x);
		for (i = 0; i < 3; ++i) g[i] = p->mu * x[i] * (x[i] * x[i] - 1.0) + 0.5 * p->lambda * (xx - 3.0) * x[i] + p->k * (x[i] - p->x0[i]);
	} break;
	default: {
		double hp = sp_dh(p, x[0] * x[1] * x[2]);
		g[0] = sp_df(p, x[0]) + sp_dg(p, x[0] * x[1]) * x[1] + sp_dg(p, x[2] * x[0]) * x[2] + hp * x[1] * x[2] + p->k * (x[0] - p->x0[0]);
		g[1] = sp_df(p, x[1]) + sp_dg(p, x[1] * x[2]) * x[2] + sp_dg(p, x[0] * x[1]) * x[0] + hp * x[2] * x[0] + p->k * (x[1] - p->x0[1]);
		g[2] = sp_df(p, x[2]) + sp_dg(p, x[2] * x[0]) * x[0] + sp_dg(p, x[1] * x[2]) * x[1] + hp * x[0] * x[1] + p->k * (x[2] - p->x0[2]);
	} break;
	}
	return prox_value(p, x);
}

/* HyperElasticTet::Prox::converged (src/TetEnergyTerm.hpp:93-95) */
static int prox_converged(const double *x0, const double *x1, const double *grad)
{
	double d[3] = {x0[0] - x1[0], x0[1] - x1[1], x0[2] - x1[2]};
	return norm3(grad) < 1e-6 || norm3(d) < 1e-6;
}

/* BacktrackingCubic::cubic (deps/mcloptlib/include/MCL/Backtracking.hpp:129-143) */
static double ls_cubic(double fx0, double gtp, double fxa, double alpha, double fxp, double alphap)
{
	double mult = 1.0 / (alpha * alpha * alphap * alphap * (alpha - alphap));
	double A00 = alphap * alphap, A01 = -alpha * alpha, A10 = -alphap * alphap * alphap, A11 = alpha * alpha * alpha;
	double B0 = fxa - fx0 - alpha * gtp, B1 = fxp - fx0 - alphap * gtp;
	double r0 = mult * (A00 * B0 + A01 * B1), r1 = mult * (A10 * B0 + A11 * B1);
	double d;
	if (fabs(r0) <= 0.0) return -gtp / (2.0 * r1);
	d = sqrt(r1 * r1 - 3.0 * r0 * gtp);
	return (-r1 + d) / (3.0 * r0);
}

/* BacktrackingCubic::search (Backtracking.hpp:79-113) with Minimizer::Settings defaults
 * ls_max_iters=100000, ls_decrease=1e-4 (Minimizer.hpp:66-70) */
static double ls_search(prox_problem *p, const double *x, const double *dir, double alpha0)
{
	const int max_iters = 100000; const double decrease = 1e-4;
	double grad[3], fx0, gtp, fxp, alphap, alpha = alpha0;
	int iter;
	if (norm3(dir) <= DBL_EPSILON) return decrease;
	fx0 = prox_gradient(p, x, grad);
	gtp = dot3(grad, dir);
	fxp = fx0; alphap = alpha;
	for (iter = 0; iter < max_iters; ++iter) {
		double xa[3] = {x[0] + alpha * dir[0], x[1] + alpha * dir[1], x[2] + alpha * dir[2]};
		double fxa = prox_value(p, xa), alpha_tmp;
		if (fxa <= fx0 + alpha * decrease * gtp) break;
		alpha_tmp = iter == 0 ? (gtp / (2.0 * (fx0 + gtp - fxa))) : ls_cubic(fx0, gtp, fxa, alpha, fxp, alphap);
		fxp = fxa; alphap = alpha;
		{ double lo = 0.1 * alpha, hi = 0.5 * alpha; alpha = alpha_tmp < lo ? lo : (alpha_tmp > hi ? hi : alpha_tmp); } /* range(): NaN compares false -> stays */
		if (alpha_tmp != alpha_tmp) alpha = alpha_tmp; /* range() returns NaN unchanged */
	}
	if (iter >= max_iters) return -1;
	return alpha;
}

/* LBFGS<double,3,8>::minimize (deps/mcloptlib/include/MCL/LBFGS.hpp:52-152), max_iters = 50 */
static int lbfgs_minimize(prox_problem *p, double *x)
{
	enum { M = 8 };
	double s[M][3], y[M][3], alpha[M], rho[M];
	double grad[3], q[3], grad_old[3], x_old[3], x_last[3];
	double gamma_k = 1.0, alpha_init = 1.0;
	int global_iter = 0, max_iters = 50, k, i, j;
	memset(s, 0, sizeof(s)); memset(y, 0, sizeof(y));
	prox_gradient(p, x, grad);
	if (p->error) return -1;
	for (k = 0; k < max_iters; ++k) {
		int iter = k < M ? k : M;
		double dir, rate, negq[3];
		memcpy(x_old, x, sizeof(x_old)); memcpy(grad_old, grad, sizeof(grad_old)); memcpy(q, grad, sizeof(q));
		global_iter++;
		for (i = iter - 1; i >= 0; --i) {
			rho[i] = 1.0 / dot3(s[i], y[i]);
			alpha[i] = rho[i] * dot3(s[i], q);
			for (j = 0; j < 3; ++j) q[j] -= alpha[i] * y[i][j];
		}
		for (j = 0; j < 3; ++j) q[j] *= gamma_k;
		for (i = 0; i < iter; ++i) {
			double beta = rho[i] * dot3(q, y[i]);
			for (j = 0; j < 3; ++j) q[j] += (alpha[i] - beta) * s[i][j];
		}
		dir = dot3(q, grad);
		if (dir <= 0) {
			double inf = fmax(fabs(grad[0]), fmax(fabs(grad[1]), fabs(grad[2])));
			memcpy(q, grad, sizeof(q));
			max_iters -= k; k = 0;
			alpha_init = fmin(1.0, 1.0 / inf);
		}
		for (j = 0; j < 3; ++j) negq[j] = -q[j];
		rate = ls_search(p, x, negq, alpha_init);
		if (p->error) return -1;
		if (rate <= 0) return -1; /* Minimizer::FAILURE */
		memcpy(x_last, x, sizeof(x_last));
		for (j = 0; j < 3; ++j) x[j] -= rate * q[j];
		if (prox_converged(x_last, x, grad)) break;
		prox_gradient(p, x, grad);
		if (p->error) return -1;
		{
			double st[3], yt[3], denom;
			for (j = 0; j < 3; ++j) { st[j] = x[j] - x_old[j]; yt[j] = grad[j] - grad_old[j]; }
			if (k < M) { memcpy(s[k], st, sizeof(st)); memcpy(y[k], yt, sizeof(yt)); }
			else {
				for (i = 0; i < M - 1; ++i) { memcpy(s[i], s[i + 1], sizeof(st)); memcpy(y[i], y[i + 1], sizeof(yt)); }
				memcpy(s[M - 1], st, sizeof(st)); memcpy(y[M - 1], yt, sizeof(yt));
			}
			denom = dot3(yt, yt);
			if (fabs(denom) <= 0) break;
			gamma_k = dot3(st, yt) / denom;
			alpha_init = 1.0;
		}
	}
	return global_iter;
}

/* TetEnergyTerm::prox (src/TetEnergyTerm.cpp:73-92) */
static void prox_tet_linear(double *z)
{
	double U[9], S[3], V[9], one[3] = {1, 1, 1}, P[9];
	int i;
	jacobi_svd(3, 3, z, U, S, V);
	if (det3cm(z) < 0.0) one[2] = -1.0;
	usvt(U, one, V, P);
	for (i = 0; i < 9; ++i) z[i] = 0.5 * (P[i] + z[i]);
}

/* HyperElasticTet::prox (src/TetEnergyTerm.cpp:114-136) */
static int prox_tet_hyper(prox_problem *p, double *z)
{
	double S[3], U[9], V[9];
	const double eps = 1e-6;
	signed_svd(z, S, U, V);
	memcpy(p->x0, S, sizeof(S));
	p->error = 0;
	if (fabs(S[0]) < eps && fabs(S[1]) < eps && fabs(S[2]) < eps) { S[0] = eps; S[1] = eps; S[2] = eps; }
	if (S[2] < 0.0) S[2] = -S[2];
	lbfgs_minimize(p, S);
	usvt(U, S, V, z);
	return p->error;
}

/* TriEnergyTerm::prox (src/TriEnergyTerm.cpp:73-101) */
static void prox_tri(double limit_min, double limit_max, double *z)
{
	double U[6], S[2], V[4], P[6];
	int r, c;
	jacobi_svd(3, 2, z, U, S, V);
	for (c = 0; c < 2; ++c) for (r = 0; r < 3; ++r) P[3 * c + r] = U[r] * V[c] + U[3 + r] * V[2 + c];
	for (r = 0; r < 6; ++r) z[r] = 0.5 * (P[r] + z[r]);
	if (limit_min > 0.0 || limit_max < 99.0) {
		double l0 = norm3(z), l1 = norm3(z + 3);
		if (l0 < limit_min) for (r = 0; r < 3; ++r) z[r] *= limit_min / l0;
		if (l1 < limit_min) for (r = 0; r < 3; ++r) z[3 + r] *= limit_min / l1;
		if (l0 > limit_max) for (r = 0; r < 3; ++r) z[r] *= limit_max / l0;
		if (l1 > limit_max) for (r = 0; r < 3; ++r) z[3 + r] *= limit_max / l1;
	}
}

/* ------------------------------------------------------------------------------------------- */
/* solver state                                                                                 */
/* ------------------------------------------------------------------------------------------- */
typedef struct {
	int kind;          /* 0 tet, 1 tri, 2 pin */
	int idx[4];
	double binv[9];    /* tet: edges_inv(c,r) at [3c+r]; tri: rest_pose(c,r) at [2c+r] */
	double weight;
	int g_index, dim;
	int model; double mu, lambda, kappa, k, limit_min, limit_max;
	double pin[3]; int active;
} term_t;

typedef struct { int kind; double p[4]; } obstacle_t;

typedef struct { int n; int *rowptr, *cols; double *vals; } csr_t;

typedef struct oracle_solver {
	int n_nodes;
	double *x, *v, *m;
	term_t *terms; int n_terms, cap_terms;
	int *pin_idx; double *pin_pos; int n_pins;
	obstacle_t obs[8]; int n_obs;
	int n_colors; int *color_off, *color_nodes;
	double dt, gravity; int admm_iters, linsolver;
	int gs_max_iters; double gs_tol, gs_omega;
	int uz_max_iters; double uz_tol;
	int n_rows;
	csr_t A;           /* 3n x 3n, = M + dt^2 D^T W^2 D (src/Solver.cpp:226) */
	/* LDL^T of A, natural ordering */
	int *Lp, *Li; double *Lx, *Dg; int have_ldlt;
	double *uz_y; int uz_rows;
	int *surf; int n_surf;   /* Solver::surface_inds (src/Solver.hpp:69): the vertices Collider::detect tests, in order; none = all */
	double uz_ck;            /* sqrt(max(0, constraint_w)), ConstraintSet::make_matrix (src/ConstraintSet.hpp:66) */
	double global_ms, local_ms; int inner_iters;
	/* Solver::ext_forces (src/Solver.hpp:71): WindForce objects, applied in order at the top of step() (src/Solver.cpp:53-54) */
	struct { int *tris; int n_tris; double dir[3]; } wind[4]; int n_wind, wind_sequential;
	int initialized;
	char err[256];
} oracle_solver;

static void csr_free(csr_t *a) { free(a->rowptr); free(a->cols); free(a->vals); memset(a, 0, sizeof(*a)); }

/* rows of D and their values for one term: EnergyTerm::get_reduction
 * (src/TetEnergyTerm.cpp:50-71, src/TriEnergyTerm.cpp:54-70, src/SpringEnergyTerm.hpp:54-59).
 * Calls emit(row_local, col, value). */
typedef void (*emit_fn)(void *ctx, int row, int col, double val);
static void term_reduction(const term_t *t, emit_fn emit, void *ctx)
{
	int r, c, j;
	if (t->kind == 0) {
		double D[4][3];
		for (r = 0; r < 3; ++r) { D[0][r] = -(t->binv[r] + t->binv[3 + r] + t->binv[6 + r]); for (c = 0; c < 3; ++c) D[c + 1][r] = t->binv[3 * c + r]; }
		for (r = 0; r < 3; ++r) for (c = 0; c < 4; ++c) for (j = 0; j < 3; ++j) emit(ctx, 3 * r + j, 3 * t->idx[c] + j, D[c][r]);
	} else if (t->kind == 1) {
		double D[3][2];
		for (r = 0; r < 2; ++r) { D[0][r] = -(t->binv[r] + t->binv[2 + r]); D[1][r] = t->binv[r]; D[2][r] = t->binv[2 + r]; }
		for (j = 0; j < 3; ++j) for (c = 0; c < 3; ++c) { emit(ctx, j, 3 * t->idx[c] + j, D[c][0]); emit(ctx, 3 + j, 3 * t->idx[c] + j, D[c][1]); }
	} else {
		for (j = 0; j < 3; ++j) emit(ctx, j, 3 * t->idx[0] + j, 1.0);
	}
}

/* D_i x for one term */
typedef struct { const double *x; double *out; } dix_ctx;
static void emit_dix(void *c_, int row, int col, double val) { dix_ctx *c = (dix_ctx *)c_; c->out[row] += val * c->x[col]; }

/* adds dt^2 w^2 D_i^T y to b */
typedef struct { const double *y; double *b; double s; } dty_ctx;
static void emit_dty(void *c_, int row, int col, double val) { dty_ctx *c = (dty_ctx *)c_; c->b[col] += c->s * val * c->y[row]; }

/* EnergyTerm::update (src/EnergyTerm.hpp:130-140) */
static int term_update(const term_t *t, const double *x, double *z, double *u)
{
	double dix[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, zi[9];
	dix_ctx c; int i, rc = 0;
	c.x = x; c.out = dix;
	term_reduction(t, emit_dix, &c);
	for (i = 0; i < t->dim; ++i) zi[i] = dix[i] + u[t->g_index + i];
	if (t->kind == 0) {
		if (t->model == TET_LINEAR) prox_tet_linear(zi);
		else { prox_problem p; memset(&p, 0, sizeof(p)); p.model = t->model; p.mu = t->mu; p.lambda = t->lambda; p.kappa = t->kappa; p.k = t->k; rc = prox_tet_hyper(&p, zi); }
	} else if (t->kind == 1) prox_tri(t->limit_min, t->limit_max, zi);
	else { if (t->active) { zi[0] = t->pin[0]; zi[1] = t->pin[1]; zi[2] = t->pin[2]; } /* rows 3..5 are dead */ }
	for (i = 0; i < t->dim; ++i) { u[t->g_index + i] += dix[i] - zi[i]; z[t->g_index + i] = zi[i]; }
	return rc;
}

/* ------------------------------------------------------------------------------------------- */
/* A = M + dt^2 D^T W^2 D  (src/Solver.cpp:207-226)                                             */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int row, col; double val; } trip_t;
static int trip_cmp(const void *a_, const void *b_) { const trip_t *a = (const trip_t *)a_, *b = (const trip_t *)b_; if (a->row != b->row) return a->row < b->row ? -1 : 1; if (a->col != b->col) return a->col < b->col ? -1 : 1; return 0; }
typedef struct { int rows[64], cols[64]; double vals[64]; int n; } termtrip_ctx;
static void emit_collect(void *c_, int row, int col, double val) { termtrip_ctx *c = (termtrip_ctx *)c_; c->rows[c->n] = row; c->cols[c->n] = col; c->vals[c->n] = val; c->n++; }

static int build_A(oracle_solver *s)
{
	size_t cap = (size_t)s->n_terms * 144 + (size_t)3 * s->n_nodes, n = 0, i;
	trip_t *tr = (trip_t *)malloc(cap * sizeof(trip_t));
	int t, a, b, dof = 3 * s->n_nodes;
	const double dt2 = s->dt * s->dt;
	if (!tr) return 1;
	for (t = 0; t < s->n_terms; ++t) {
		termtrip_ctx c; double w2 = dt2 * s->terms[t].weight * s->terms[t].weight;
		c.n = 0;
		term_reduction(&s->terms[t], emit_collect, &c);
		for (a = 0; a < c.n; ++a) for (b = 0; b < c.n; ++b) if (c.rows[a] == c.rows[b]) { tr[n].row = c.cols[a]; tr[n].col = c.cols[b]; tr[n].val = w2 * c.vals[a] * c.vals[b]; n++; }
	}
	for (t = 0; t < dof; ++t) { tr[n].row = t; tr[n].col = t; tr[n].val = s->m[t]; n++; }
	qsort(tr, n, sizeof(trip_t), trip_cmp);
	csr_free(&s->A);
	s->A.n = dof;
	s->A.rowptr = (int *)calloc(dof + 1, sizeof(int));
	s->A.cols = (int *)malloc(n * sizeof(int));
	s->A.vals = (double *)malloc(n * sizeof(double));
	{
		size_t k = 0;
		for (i = 0; i < n;) {
			size_t j = i; double sum = 0;
			while (j < n && tr[j].row == tr[i].row && tr[j].col == tr[i].col) { sum += tr[j].val; ++j; }
			s->A.cols[k] = tr[i].col; s->A.vals[k] = sum; s->A.rowptr[tr[i].row + 1]++; ++k;
			i = j;
		}
	}
	for (t = 0; t < dof; ++t) s->A.rowptr[t + 1] += s->A.rowptr[t];
	free(tr);
	return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* LDL^T, natural ordering (stands in for Eigen::SimplicialLDLT, src/LinearSolver.hpp:79-90)    */
/* ------------------------------------------------------------------------------------------- */
static int factor_ldlt(oracle_solver *s)
{
	const csr_t *A = &s->A; int n = A->n, k, p;
	int *parent = (int *)malloc(n * sizeof(int)), *flag = (int *)malloc(n * sizeof(int)), *lnz = (int *)calloc(n, sizeof(int)), *pattern = (int *)malloc(n * sizeof(int));
	double *Y = (double *)calloc(n, sizeof(double));
	free(s->Lp); free(s->Li); free(s->Lx); free(s->Dg);
	s->Lp = (int *)calloc(n + 1, sizeof(int)); s->Dg = (double *)calloc(n, sizeof(double));
	/* row k of the symmetric CSR restricted to cols <= k is column k of the upper triangle */
	for (k = 0; k < n; ++k) {
		parent[k] = -1; flag[k] = k;
		for (p = A->rowptr[k]; p < A->rowptr[k + 1]; ++p) { int i = A->cols[p]; if (i < k) for (; flag[i] != k; i = parent[i]) { if (parent[i] == -1) parent[i] = k; lnz[i]++; flag[i] = k; } }
	}
	for (k = 0; k < n; ++k) s->Lp[k + 1] = s->Lp[k] + lnz[k];
	s->Li = (int *)malloc((size_t)(s->Lp[n] > 0 ? s->Lp[n] : 1) * sizeof(int)); s->Lx = (double *)malloc((size_t)(s->Lp[n] > 0 ? s->Lp[n] : 1) * sizeof(double));
	memset(lnz, 0, n * sizeof(int));
	for (k = 0; k < n; ++k) {
		int top = n;
		flag[k] = k;
		for (p = A->rowptr[k]; p < A->rowptr[k + 1]; ++p) {
			int i = A->cols[p], len = 0;
			if (i > k) continue;
			Y[i] += A->vals[p];
			for (; flag[i] != k; i = parent[i]) { pattern[len++] = i; flag[i] = k; }
			while (len > 0) pattern[--top] = pattern[--len];
		}
		s->Dg[k] = Y[k]; Y[k] = 0.0;
		for (; top < n; ++top) {
			int i = pattern[top], p2 = s->Lp[i] + lnz[i], q; double yi = Y[i], lki;
			Y[i] = 0.0;
			for (q = s->Lp[i]; q < p2; ++q) Y[s->Li[q]] -= s->Lx[q] * yi;
			lki = yi / s->Dg[i];
			s->Dg[k] -= lki * yi;
			s->Li[p2] = k; s->Lx[p2] = lki; lnz[i]++;
		}
		if (s->Dg[k] == 0.0) { free(parent); free(flag); free(lnz); free(pattern); free(Y); return 1; }
	}
	free(parent); free(flag); free(lnz); free(pattern); free(Y);
	s->have_ldlt = 1;
	return 0;
}

static void ldlt_solve(const oracle_solver *s, const double *b, double *x)
{
	int n = s->A.n, j, p;
	if (x != b) memcpy(x, b, n * sizeof(double));
	for (j = 0; j < n; ++j) for (p = s->Lp[j]; p < s->Lp[j + 1]; ++p) x[s->Li[p]] -= s->Lx[p] * x[j];
	for (j = 0; j < n; ++j) x[j] /= s->Dg[j];
	for (j = n - 1; j >= 0; --j) for (p = s->Lp[j]; p < s->Lp[j + 1]; ++p) x[j] -= s->Lx[p] * x[s->Li[p]];
}

/* ------------------------------------------------------------------------------------------- */
/* passive obstacles and NodalMultiColorGS                                                      */
/* ------------------------------------------------------------------------------------------- */
/* Collider::detect_passive (src/Collider.hpp:137-150), Floor/Sphere::signed_distance
 * (src/PassiveObject.hpp:32-64) */
static int detect_passive(const oracle_solver *s, const double *x, double *n, double *p)
{
	double dx = DBL_MAX; int j;
	for (j = 0; j < s->n_obs; ++j) {
		const obstacle_t *o = &s->obs[j];
		if (o->kind == 0) {
			double d = x[1] - o->p[0];
			if (!(d > dx)) { dx = d; p[0] = x[0]; p[1] = o->p[0]; p[2] = x[2]; n[0] = 0; n[1] = 1; n[2] = 0; }
		} else {
			double dir[3] = {x[0] - o->p[0], x[1] - o->p[1], x[2] - o->p[2]}, len = norm3(dir), d = len - o->p[3];
			if (!(d > dx)) { int k; dx = d; for (k = 0; k < 3; ++k) { dir[k] /= len; p[k] = o->p[k] + dir[k] * o->p[3]; n[k] = dir[k]; } }
		}
		if (dx < 0) return 1;
	}
	return 0;
}

/* NodalMultiColorGS::solve (src/NodalMultiColorGS.hpp:60-146) with segment_update (:180-215),
 * constrained_segment_update (:218-262) and orthoG (:171-177).  No dynamic collisions (C empty). */
static int mcgs_solve(oracle_solver *s, double *x, const double *b)
{
	const csr_t *A = &s->A; int dof = A->n, iter, color, i, sx, p;
	double b_norm = 1.0, tol2 = s->gs_tol * s->gs_tol;
	int *pin_of = NULL;
	if (s->gs_tol > 0) { b_norm = 0; for (i = 0; i < dof; ++i) b_norm += b[i] * b[i]; }
	if (s->n_pins > 0) { pin_of = (int *)malloc(s->n_nodes * sizeof(int)); for (i = 0; i < s->n_nodes; ++i) pin_of[i] = -1; for (i = 0; i < s->n_pins; ++i) pin_of[s->pin_idx[i]] = i; }
	for (iter = 0; iter < s->gs_max_iters; ++iter) {
		for (color = 0; color < s->n_colors; ++color) {
			int k0 = s->color_off[color], k1 = s->color_off[color + 1], k;
			#pragma omp parallel for if (k1 - k0 > 31) private(sx, p)
			for (k = k0; k < k1; ++k) {
				int idx = s->color_nodes[k], idx3 = 3 * idx;
				double gs[3], nx[3], nrm[3], pt[3];
				if (pin_of && pin_of[idx] >= 0) { const double *pp = s->pin_pos + 3 * pin_of[idx]; x[idx3] = pp[0]; x[idx3 + 1] = pp[1]; x[idx3 + 2] = pp[2]; continue; }
				for (sx = 0; sx < 3; ++sx) {
					double LUx = 0.0, aii = 0.0;
					for (p = A->rowptr[idx3 + sx]; p < A->rowptr[idx3 + sx + 1]; ++p) {
						int c = A->cols[p];
						if (fabs(A->vals[p]) <= 0.0) continue;
						if (c == idx3 + sx) { aii = A->vals[p]; continue; }
						LUx += A->vals[p] * x[c];
					}
					gs[sx] = (b[idx3 + sx] - LUx) / aii;
					nx[sx] = (1.0 - s->gs_omega) * x[idx3 + sx] + s->gs_omega * gs[sx];
				}
				if (s->n_obs > 0 && detect_passive(s, nx, nrm, pt)) {
					double not_n[3] = {nrm[0] > 0.999 ? 0.0 : 1.0, 0.0, nrm[0] > 0.999 ? 1.0 : 0.0}, gu[3], gv[3], d[3], t0, t1, l;
					cross3(not_n, nrm, gu); l = norm3(gu); gu[0] /= l; gu[1] /= l; gu[2] /= l;
					cross3(nrm, gu, gv); l = norm3(gv); gv[0] /= l; gv[1] /= l; gv[2] /= l;
					d[0] = gs[0] - pt[0]; d[1] = gs[1] - pt[1]; d[2] = gs[2] - pt[2];
					t0 = dot3(gu, d); t1 = dot3(gv, d);
					nx[0] = (gu[0] * t0 + gv[0] * t1) + pt[0]; nx[1] = (gu[1] * t0 + gv[1] * t1) + pt[1]; nx[2] = (gu[2] * t0 + gv[2] * t1) + pt[2];
				}
				x[idx3] = nx[0]; x[idx3 + 1] = nx[1]; x[idx3 + 2] = nx[2];
			}
		}
		if (s->gs_tol > 0) {
			double err2 = 0;
			for (i = 0; i < dof; ++i) { double r = b[i]; for (p = A->rowptr[i]; p < A->rowptr[i + 1]; ++p) r -= A->vals[p] * x[A->cols[p]]; err2 += r * r; }
			if (err2 / b_norm < tol2) break;
		}
	}
	free(pin_of);
	return iter;
}

/* UzawaCG::solve (src/UzawaCG.hpp:57-125) with ConstraintSet::make_matrix (src/ConstraintSet.hpp:59-116)
 * for passive hits only (Collider::detect over Solver::surface_inds with_passive, src/Collider.hpp:152-212); rows are
 * scaled by ck = sqrt(constraint_w), constraint_w = 1 unless -ck overrides it (src/Solver.cpp:239,245). */
static int uzawa_solve(oracle_solver *s, double *x, const double *b, const double *curr_x)
{
	int dof = s->A.n, n = s->n_nodes, i, rows = 0, iter, c;
	const int n_cand = s->n_surf > 0 ? s->n_surf : n;
	const double ck = s->uz_ck;
	int *hv = (int *)malloc((size_t)(n > s->n_surf ? n : s->n_surf) * sizeof(int));
	double *hn = (double *)malloc((size_t)3 * (n > s->n_surf ? n : s->n_surf) * sizeof(double)), *hc = (double *)malloc((size_t)(n > s->n_surf ? n : s->n_surf) * sizeof(double));
	double *q1, *q2, *r, *d, *q3;
	const double tol2 = s->uz_tol * s->uz_tol;
	for (c = 0; c < n_cand && s->n_obs > 0; ++c) {
		/* Collider::detect: every passive object lowers the payload, hit if dx < 0 */
		i = s->n_surf > 0 ? s->surf[c] : c;
		double dx = DBL_MAX, nn[3] = {0, 0, 0}, pp[3] = {0, 0, 0}; int j;
		const double *xi = curr_x + 3 * i;
		for (j = 0; j < s->n_obs; ++j) {
			const obstacle_t *o = &s->obs[j];
			if (o->kind == 0) { double dd = xi[1] - o->p[0]; if (!(dd > dx)) { dx = dd; pp[0] = xi[0]; pp[1] = o->p[0]; pp[2] = xi[2]; nn[0] = 0; nn[1] = 1; nn[2] = 0; } }
			else { double dir[3] = {xi[0] - o->p[0], xi[1] - o->p[1], xi[2] - o->p[2]}, len = norm3(dir), dd = len - o->p[3]; if (!(dd > dx)) { int k; dx = dd; for (k = 0; k < 3; ++k) { dir[k] /= len; pp[k] = o->p[k] + dir[k] * o->p[3]; nn[k] = dir[k]; } } }
		}
		if (dx < 0) { hv[rows] = i; hn[3 * rows] = ck * nn[0]; hn[3 * rows + 1] = ck * nn[1]; hn[3 * rows + 2] = ck * nn[2]; hc[rows] = ck * dot3(nn, pp); rows++; }
	}
	if (s->uz_rows != rows) { free(s->uz_y); s->uz_y = (double *)calloc(rows > 0 ? rows : 1, sizeof(double)); s->uz_rows = rows; }
	if (rows == 0) { ldlt_solve(s, b, x); free(hv); free(hn); free(hc); return 1; }
	q1 = (double *)malloc(dof * sizeof(double)); q2 = (double *)malloc(dof * sizeof(double));
	r = (double *)malloc(rows * sizeof(double)); d = (double *)malloc(rows * sizeof(double)); q3 = (double *)malloc(rows * sizeof(double));
	memcpy(q1, b, dof * sizeof(double));
	for (i = 0; i < rows; ++i) { int k; for (k = 0; k < 3; ++k) q1[3 * hv[i] + k] -= hn[3 * i + k] * s->uz_y[i]; }
	ldlt_solve(s, q1, x);
	for (i = 0; i < rows; ++i) { r[i] = dot3(hn + 3 * i, x + 3 * hv[i]) - hc[i]; d[i] = r[i]; }
	for (iter = 0; iter < s->uz_max_iters; ++iter) {
		double denom = 0, dr = 0, alpha, beta, rr = 0, rq = 0;
		memset(q1, 0, dof * sizeof(double));
		for (i = 0; i < rows; ++i) { int k; for (k = 0; k < 3; ++k) q1[3 * hv[i] + k] += hn[3 * i + k] * d[i]; }
		ldlt_solve(s, q1, q2);
		for (i = 0; i < rows; ++i) { q3[i] = dot3(hn + 3 * i, q2 + 3 * hv[i]); denom += d[i] * q3[i]; dr += d[i] * r[i]; }
		if (fabs(denom) < DBL_MIN) break;
		alpha = dr / denom;
		for (i = 0; i < dof; ++i) x[i] -= alpha * q2[i];
		for (i = 0; i < rows; ++i) { s->uz_y[i] += alpha * d[i]; r[i] -= alpha * q3[i]; rr += r[i] * r[i]; }
		if (rr < tol2) break;
		denom = 0; for (i = 0; i < rows; ++i) { denom += d[i] * q3[i]; rq += r[i] * q3[i]; }
		if (fabs(denom) < DBL_MIN) break;
		beta = rq / denom;
		for (i = 0; i < rows; ++i) d[i] = r[i] - beta * d[i];
	}
	free(q1); free(q2); free(r); free(d); free(q3); free(hv); free(hn); free(hc);
	return iter;
}

static int linsolve(oracle_solver *s, double *x, const double *b)
{
	if (s->linsolver == 1) return mcgs_solve(s, x, b);
	if (s->linsolver == 2) { double *cx = (double *)malloc(s->A.n * sizeof(double)); int it; memcpy(cx, x, s->A.n * sizeof(double)); it = uzawa_solve(s, x, b, cx); free(cx); return it; }
	ldlt_solve(s, b, x);
	return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* public API (ctypes)                                                                          */
/* ------------------------------------------------------------------------------------------- */
oracle_solver *oracle_create(void)
{
	oracle_solver *s = (oracle_solver *)calloc(1, sizeof(oracle_solver));
	s->gs_max_iters = 30; s->gs_tol = 1e-10; s->gs_omega = 1.9; /* src/NodalMultiColorGS.hpp:45-46 */
	s->uz_max_iters = 20; s->uz_tol = 1e-10;                     /* src/UzawaCG.hpp:45-46 */
	s->uz_ck = 1.0; s->surf = NULL; s->n_surf = 0;
	return s;
}
void oracle_destroy(oracle_solver *s)
{
	if (!s) return;
	free(s->x); free(s->v); free(s->m); free(s->terms); free(s->pin_idx); free(s->pin_pos); free(s->color_off); free(s->color_nodes);
	{ int w; for (w = 0; w < s->n_wind; ++w) free(s->wind[w].tris); }
	csr_free(&s->A); free(s->Lp); free(s->Li); free(s->Lx); free(s->Dg); free(s->uz_y); free(s);
}
const char *oracle_last_error(const oracle_solver *s) { return s->err; }

/* Solver::add_nodes (src/Solver.hpp:127-141) */
int oracle_add_nodes(oracle_solver *s, const double *x, const double *m, int n_verts)
{
	int prev = s->n_nodes, tot = prev + n_verts;
	s->x = (double *)realloc(s->x, (size_t)3 * tot * sizeof(double)); s->v = (double *)realloc(s->v, (size_t)3 * tot * sizeof(double)); s->m = (double *)realloc(s->m, (size_t)3 * tot * sizeof(double));
	memcpy(s->x + 3 * prev, x, (size_t)3 * n_verts * sizeof(double)); memcpy(s->m + 3 * prev, m, (size_t)3 * n_verts * sizeof(double));
	memset(s->v + 3 * prev, 0, (size_t)3 * n_verts * sizeof(double));
	s->n_nodes = tot;
	return tot;
}

static term_t *new_term(oracle_solver *s)
{
	if (s->n_terms == s->cap_terms) { s->cap_terms = s->cap_terms ? 2 * s->cap_terms : 1024; s->terms = (term_t *)realloc(s->terms, (size_t)s->cap_terms * sizeof(term_t)); }
	memset(&s->terms[s->n_terms], 0, sizeof(term_t));
	return &s->terms[s->n_terms++];
}

/* create_tets_from_mesh + TetEnergyTerm::TetEnergyTerm (src/TetEnergyTerm.hpp:35-51, src/TetEnergyTerm.cpp:31-48) */
int oracle_add_tets(oracle_solver *s, const double *verts, const int *inds, int n_tets, int model, double mu, double lambda, double kappa, int vertex_offset)
{
	int i, c, r;
	for (i = 0; i < n_tets; ++i) {
		term_t *t = new_term(s);
		double e[9], det, id, vol;
		const double *v0 = verts + 3 * inds[4 * i];
		for (c = 0; c < 3; ++c) { const double *vc = verts + 3 * inds[4 * i + c + 1]; for (r = 0; r < 3; ++r) e[3 * r + c] = vc[r] - v0[r]; }
		det = e[0] * (e[4] * e[8] - e[5] * e[7]) - e[1] * (e[3] * e[8] - e[5] * e[6]) + e[2] * (e[3] * e[7] - e[4] * e[6]);
		id = 1.0 / det;
		t->binv[0] = (e[4] * e[8] - e[5] * e[7]) * id; t->binv[1] = (e[2] * e[7] - e[1] * e[8]) * id; t->binv[2] = (e[1] * e[5] - e[2] * e[4]) * id;
		t->binv[3] = (e[5] * e[6] - e[3] * e[8]) * id; t->binv[4] = (e[0] * e[8] - e[2] * e[6]) * id; t->binv[5] = (e[2] * e[3] - e[0] * e[5]) * id;
		t->binv[6] = (e[3] * e[7] - e[4] * e[6]) * id; t->binv[7] = (e[1] * e[6] - e[0] * e[7]) * id; t->binv[8] = (e[0] * e[4] - e[1] * e[3]) * id;
		vol = det / 6.0f;
		if (vol < 0) { snprintf(s->err, sizeof(s->err), "**TetEnergyTerm Error: Inverted initial tet"); s->n_terms--; return 1; }
		t->kind = 0; t->dim = 9; t->model = model; t->mu = mu; t->lambda = lambda; t->kappa = kappa;
		t->k = lambda + (2.0 / 3.0) * mu; /* Lame::bulk_modulus (src/EnergyTerm.hpp:41) */
		t->weight = sqrt(t->k * vol);
		for (c = 0; c < 4; ++c) t->idx[c] = inds[4 * i + c] + vertex_offset;
	}
	return 0;
}

/* SplineTet(tet, verts, lame, spline) with a spline whose constants differ from the element's Lame
 * (src/TetEnergyTerm.hpp:200-205): the weight and the prox penalty K come from the Lame, the energy from the spline. */
int oracle_add_spline_tets(oracle_solver *s, const double *verts, const int *inds, int n_tets, int model, double mu, double lambda,
	double sp_mu, double sp_lambda, double sp_kappa, int vertex_offset)
{
	const int first = s->n_terms;
	int i, rc = oracle_add_tets(s, verts, inds, n_tets, model, mu, lambda, sp_kappa, vertex_offset);
	if (rc) return rc;
	for (i = first; i < s->n_terms; ++i) { s->terms[i].mu = sp_mu; s->terms[i].lambda = sp_lambda; } /* k and weight stay the Lame's */
	return 0;
}

/* create_tris_from_mesh + TriEnergyTerm::TriEnergyTerm (src/TriEnergyTerm.hpp:31-46, src/TriEnergyTerm.cpp:29-51) */
int oracle_add_tris(oracle_solver *s, const double *verts, const int *inds, int n_tris, double mu, double lambda, double limit_min, double limit_max, int vertex_offset)
{
	int i, c;
	if (limit_min > 1.0 || limit_max < 1.0) { snprintf(s->err, sizeof(s->err), "**TriEnergyTerm Error: bad strain limits"); return 1; }
	for (i = 0; i < n_tris; ++i) {
		term_t *t = new_term(s);
		const double *v0 = verts + 3 * inds[3 * i], *v1 = verts + 3 * inds[3 * i + 1], *v2 = verts + 3 * inds[3 * i + 2];
		double e12[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e13[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
		double n1[3], n2[3], l, d, b00, b01, b10, b11, det, area;
		l = norm3(e12); n1[0] = e12[0] / l; n1[1] = e12[1] / l; n1[2] = e12[2] / l;
		d = dot3(e13, n1); n2[0] = e13[0] - d * n1[0]; n2[1] = e13[1] - d * n1[1]; n2[2] = e13[2] - d * n1[2];
		l = norm3(n2); n2[0] /= l; n2[1] /= l; n2[2] /= l;
		b00 = dot3(n1, e12); b01 = dot3(n1, e13); b10 = dot3(n2, e12); b11 = dot3(n2, e13);
		det = b00 * b11 - b01 * b10;
		t->binv[0] = b11 / det; t->binv[1] = -b01 / det; t->binv[2] = -b10 / det; t->binv[3] = b00 / det;
		area = det / 2.0f;
		if (area < 0) { snprintf(s->err, sizeof(s->err), "**TriEnergyTerm Error: Inverted initial pose"); s->n_terms--; return 1; }
		t->kind = 1; t->dim = 6; t->mu = mu; t->lambda = lambda; t->k = lambda + (2.0 / 3.0) * mu; t->limit_min = limit_min; t->limit_max = limit_max;
		t->weight = sqrt(t->k * area);
		for (c = 0; c < 3; ++c) t->idx[c] = inds[3 * i + c] + vertex_offset;
	}
	return 0;
}

/* Solver::set_pins (src/Solver.cpp:113-157); points == NULL pins in place */
int oracle_set_pins(oracle_solver *s, const int *inds, const double *points, int n)
{
	int i, t;
	free(s->pin_idx); free(s->pin_pos);
	s->pin_idx = (int *)malloc((n > 0 ? n : 1) * sizeof(int)); s->pin_pos = (double *)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(double));
	s->n_pins = n;
	for (i = 0; i < n; ++i) { s->pin_idx[i] = inds[i]; memcpy(s->pin_pos + 3 * i, points ? points + 3 * i : s->x + 3 * inds[i], 3 * sizeof(double)); }
	if (s->initialized && (s->linsolver == 0 || s->linsolver == 2)) {
		for (t = 0; t < s->n_terms; ++t) if (s->terms[t].kind == 2) s->terms[t].active = 0;
		for (i = 0; i < n; ++i) {
			int found = 0;
			for (t = 0; t < s->n_terms; ++t) if (s->terms[t].kind == 2 && s->terms[t].idx[0] == inds[i]) { s->terms[t].active = 1; memcpy(s->terms[t].pin, s->pin_pos + 3 * i, 3 * sizeof(double)); found = 1; }
			if (!found) { snprintf(s->err, sizeof(s->err), "**Solver::set_pins Error: Constraint for %d not found.", inds[i]); return 1; }
		}
	}
	return 0;
}

int oracle_add_obstacle(oracle_solver *s, int kind, const double *params)
{
	if (s->n_obs >= 8) return 1;
	s->obs[s->n_obs].kind = kind; memcpy(s->obs[s->n_obs].p, params, 4 * sizeof(double)); s->n_obs++;
	return 0;
}

int oracle_set_colors(oracle_solver *s, int n_colors, const int *offsets, const int *nodes)
{
	free(s->color_off); free(s->color_nodes);
	s->n_colors = n_colors;
	s->color_off = (int *)malloc((n_colors + 1) * sizeof(int)); memcpy(s->color_off, offsets, (n_colors + 1) * sizeof(int));
	s->color_nodes = (int *)malloc((offsets[n_colors] > 0 ? offsets[n_colors] : 1) * sizeof(int)); memcpy(s->color_nodes, nodes, offsets[n_colors] * sizeof(int));
	return 0;
}

void oracle_gs_params(oracle_solver *s, int max_iters, double tol, double omega) { s->gs_max_iters = max_iters; s->gs_tol = tol; s->gs_omega = omega; }

/* Solver::initialize (src/Solver.cpp:167-261).  Pins become SpringPin terms for linsolver 0/2, in
 * the order they were given (the reference iterates an unordered_map; tests read g_index back). */
/* Solver::surface_inds and Settings::constraint_w (-ck) for the UzawaCG path; call before oracle_initialize */
int oracle_set_uzawa(oracle_solver *s, int n_surf, const int *surf, double constraint_w)
{
	free(s->surf); s->surf = NULL; s->n_surf = 0;
	if (n_surf > 0) { s->surf = (int *)malloc((size_t)n_surf * sizeof(int)); memcpy(s->surf, surf, (size_t)n_surf * sizeof(int)); s->n_surf = n_surf; }
	s->uz_ck = constraint_w > 0.0 ? sqrt(constraint_w) : 1.0;
	return 0;
}

int oracle_initialize(oracle_solver *s, double dt, int admm_iters, double gravity, int linsolver)
{
	int t, i, rows = 0;
	if (dt <= 0.0) dt = 1.0 / 24.0;
	if (s->n_nodes < 1) { snprintf(s->err, sizeof(s->err), "**Solver Error: Problem with node data!"); return 2; }
	s->dt = dt; s->admm_iters = admm_iters; s->gravity = gravity; s->linsolver = linsolver;
	memset(s->v, 0, (size_t)3 * s->n_nodes * sizeof(double));
	if (!s->initialized && (linsolver == 0 || linsolver == 2)) {
		for (i = 0; i < s->n_pins; ++i) {
			term_t *p = new_term(s);
			p->kind = 2; p->dim = 6; p->idx[0] = s->pin_idx[i]; p->active = 1; memcpy(p->pin, s->pin_pos + 3 * i, 3 * sizeof(double));
			{ double E = 10000000, nu = 0.499, mu = E / (2.0 * (1.0 + nu)), lambda = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)); p->weight = sqrt((lambda + (2.0 / 3.0) * mu) * 2.0); } /* src/SpringEnergyTerm.hpp:46-51 */
		}
	}
	for (t = 0; t < s->n_terms; ++t) { if (s->terms[t].weight <= 0.0) { snprintf(s->err, sizeof(s->err), "**EnergyTerm::get_reduction Error: Some weight leq 0"); return 1; } s->terms[t].g_index = rows; rows += s->terms[t].dim; }
	s->n_rows = rows;
	if (build_A(s)) return 1;
	if (linsolver == 1) { if (s->n_colors <= 0) { snprintf(s->err, sizeof(s->err), "oracle: colours must be supplied for NodalMultiColorGS"); return 1; } }
	else {
		if (linsolver == 0 && s->n_obs > 0) { snprintf(s->err, sizeof(s->err), "**Solver::add_obstacle Error: No collisions with LDLT solver"); return 1; }
		if (factor_ldlt(s)) { snprintf(s->err, sizeof(s->err), "oracle: zero pivot"); return 1; }
	}
	s->initialized = 1;
	return 0;
}

/* WindForce::project (src/ExplicitForce.cpp:47-104), Wejchert & Haumann: per triangle the mean node velocity relative to
   the wind, its component along the unit normal, force = -alpha_n area v_n |v_n| n, a 0.33 share times dt added to the
   VELOCITY of each of the three nodes (no division by mass in the reference).
   The reference loops over the triangles with `omp parallel for` and adds the shares under an `omp critical` while other
   threads are reading the same velocities (:62-66 against :96-100): with one thread every triangle sees the velocities
   already changed by the triangles before it; with several threads the outcome depends on the interleaving.
   sequential = 1 restates the one-thread order (pinned against the compiled reference run with one thread,
   tests/test_oracle_vs_ref.py); sequential = 0 forms every force from the velocities BEFORE the call, the only
   order-independent reading and the one the device implements (csrc/kernels.cuh: wind_*_kernel). */
void oracle_wind_project(int n_tris, const int *tris, const double *dir, double dt, int n_nodes, const double *x, double *v, int sequential)
{
	int i, j, a;
	double *v0 = v;
	if (!sequential) {
		v0 = (double *)malloc((size_t)3 * n_nodes * sizeof(double));
		memcpy(v0, v, (size_t)3 * n_nodes * sizeof(double));
	}
	for (i = 0; i < n_tris; ++i) {
		const int id[3] = {3 * tris[3 * i], 3 * tris[3 * i + 1], 3 * tris[3 * i + 2]};
		double vr[3], e1[3], e2[3], n[3], len, area, vn, f[3];
		for (a = 0; a < 3; ++a) vr[a] = (v0[id[0] + a] + v0[id[1] + a] + v0[id[2] + a]) / 3.0 - dir[a];
		for (a = 0; a < 3; ++a) { e1[a] = x[id[1] + a] - x[id[0] + a]; e2[a] = x[id[2] + a] - x[id[0] + a]; }
		cross3(e1, e2, n);
		len = norm3(n);
		area = 0.5 * len;
		if (len > 0) for (a = 0; a < 3; ++a) n[a] /= len; /* Eigen's normalized() leaves a zero vector as it is */
		vn = dot3(n, vr);
		for (a = 0; a < 3; ++a) { f[a] = -1000.0 * area * vn * fabs(vn) * n[a]; f[a] *= 0.33; f[a] *= dt; }
		for (j = 0; j < 3; ++j) for (a = 0; a < 3; ++a) v[id[j] + a] += f[a];
	}
	if (!sequential) free(v0);
}

void oracle_wind_mode(oracle_solver *s, int sequential) { s->wind_sequential = sequential; }
int oracle_add_wind(oracle_solver *s, const int *tris, int n_tris, const double *dir)
{
	if (s->n_wind >= 4) { snprintf(s->err, sizeof(s->err), "oracle: at most 4 wind forces"); return 1; }
	s->wind[s->n_wind].tris = (int *)malloc((size_t)3 * (n_tris > 0 ? n_tris : 1) * sizeof(int));
	memcpy(s->wind[s->n_wind].tris, tris, (size_t)3 * n_tris * sizeof(int));
	s->wind[s->n_wind].n_tris = n_tris;
	memcpy(s->wind[s->n_wind].dir, dir, 3 * sizeof(double));
	s->n_wind++;
	return 0;
}

/* Solver::step (src/Solver.cpp:35-110); optional traces hold z,u (n_rows each) and b,x (dof each) per ADMM iteration */
int oracle_step_traced(oracle_solver *s, double *zt, double *ut, double *bt, double *xt)
{
	int dof = 3 * s->n_nodes, i, it, t, R = s->n_rows, bad = 0;
	double dt = s->dt;
	double *x_bar = (double *)malloc(dof * sizeof(double)), *M_xbar = (double *)malloc(dof * sizeof(double)), *cx = (double *)malloc(dof * sizeof(double));
	double *z = (double *)calloc(R > 0 ? R : 1, sizeof(double)), *u = (double *)calloc(R > 0 ? R : 1, sizeof(double)), *b = (double *)malloc(dof * sizeof(double));
	s->global_ms = s->local_ms = 0; s->inner_iters = 0;
	for (i = 0; i < s->n_wind; ++i) oracle_wind_project(s->wind[i].n_tris, s->wind[i].tris, s->wind[i].dir, dt, s->n_nodes, s->x, s->v, s->wind_sequential);
	if (fabs(s->gravity) > 0) for (i = 0; i < s->n_nodes; ++i) s->v[3 * i + 1] += dt * s->gravity;
	for (i = 0; i < dof; ++i) { x_bar[i] = s->x[i] + dt * s->v[i]; M_xbar[i] = s->m[i] * x_bar[i]; cx[i] = x_bar[i]; }
	for (it = 0; it < s->admm_iters; ++it) {
		double t0 = omp_get_wtime(), t1;
		#pragma omp parallel for schedule(dynamic, 64) reduction(+:bad)
		for (t = 0; t < s->n_terms; ++t) bad += term_update(&s->terms[t], cx, z, u);
		t1 = omp_get_wtime(); s->local_ms += 1e3 * (t1 - t0);
		/* b = M x_bar + dt^2 D^T W^2 (z - u)   (src/Solver.cpp:98) */
		memcpy(b, M_xbar, dof * sizeof(double));
		for (t = 0; t < s->n_terms; ++t) {
			const term_t *tm = &s->terms[t]; double y[9]; dty_ctx c; int k;
			for (k = 0; k < tm->dim; ++k) y[k] = z[tm->g_index + k] - u[tm->g_index + k];
			c.y = y; c.b = b; c.s = dt * dt * tm->weight * tm->weight;
			term_reduction(tm, emit_dty, &c);
		}
		s->inner_iters += linsolve(s, cx, b);
		s->global_ms += 1e3 * (omp_get_wtime() - t1);
		if (zt) memcpy(zt + (size_t)it * R, z, R * sizeof(double));
		if (ut) memcpy(ut + (size_t)it * R, u, R * sizeof(double));
		if (bt) memcpy(bt + (size_t)it * dof, b, dof * sizeof(double));
		if (xt) memcpy(xt + (size_t)it * dof, cx, dof * sizeof(double));
	}
	for (i = 0; i < dof; ++i) { s->v[i] = (cx[i] - s->x[i]) * (1.0 / dt); s->x[i] = cx[i]; }
	free(x_bar); free(M_xbar); free(cx); free(z); free(u); free(b);
	return bad ? 3 : 0;
}
int oracle_step(oracle_solver *s) { return oracle_step_traced(s, NULL, NULL, NULL, NULL); }

int oracle_dof(const oracle_solver *s) { return 3 * s->n_nodes; }
int oracle_n_rows(const oracle_solver *s) { return s->n_rows; }
int oracle_n_terms(const oracle_solver *s) { return s->n_terms; }
void oracle_get_x(const oracle_solver *s, double *x) { memcpy(x, s->x, (size_t)3 * s->n_nodes * sizeof(double)); }
void oracle_get_v(const oracle_solver *s, double *v) { memcpy(v, s->v, (size_t)3 * s->n_nodes * sizeof(double)); }
void oracle_set_x(oracle_solver *s, const double *x) { memcpy(s->x, x, (size_t)3 * s->n_nodes * sizeof(double)); }
void oracle_set_v(oracle_solver *s, const double *v) { memcpy(s->v, v, (size_t)3 * s->n_nodes * sizeof(double)); }
void oracle_set_admm_iters(oracle_solver *s, int it) { s->admm_iters = it; }
void oracle_runtime(const oracle_solver *s, double *out) { out[0] = s->global_ms; out[1] = s->local_ms; out[2] = 0; out[3] = s->inner_iters; }
void oracle_get_row_offsets(const oracle_solver *s, int *out) { int t; for (t = 0; t < s->n_terms; ++t) out[t] = s->terms[t].g_index; }
void oracle_get_weights(const oracle_solver *s, double *out) { int t; for (t = 0; t < s->n_terms; ++t) out[t] = s->terms[t].weight; }
void oracle_A_shape(const oracle_solver *s, long long *out) { out[0] = s->A.n; out[1] = s->A.rowptr ? s->A.rowptr[s->A.n] : 0; }
void oracle_A_get(const oracle_solver *s, int *rowptr, int *cols, double *vals)
{
	memcpy(rowptr, s->A.rowptr, (s->A.n + 1) * sizeof(int)); memcpy(cols, s->A.cols, s->A.rowptr[s->A.n] * sizeof(int)); memcpy(vals, s->A.vals, s->A.rowptr[s->A.n] * sizeof(double));
}
int oracle_linsolve(oracle_solver *s, double *x, const double *b) { return linsolve(s, x, b); }

/* D_i x of every term on positions x (rows laid out by g_index): the F-layout known answer of
 * samples/tests/test_lineartet.cpp:136-156 */
void oracle_apply_D(const oracle_solver *s, const double *x, double *out)
{
	int t; memset(out, 0, (size_t)s->n_rows * sizeof(double));
	for (t = 0; t < s->n_terms; ++t) { dix_ctx c; c.x = x; c.out = out + s->terms[t].g_index; term_reduction(&s->terms[t], emit_dix, &c); }
}

/* TetEnergyTerm::energy / HyperElasticTet::energy (src/TetEnergyTerm.cpp:94-101, 138-150) of term t on positions x */
double oracle_term_energy(const oracle_solver *s, int t, const double *x)
{
	const term_t *tm = &s->terms[t]; double F[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, U[9], S[3], V[9], vol; dix_ctx c;
	c.x = x; c.out = F; term_reduction(tm, emit_dix, &c);
	vol = tm->weight * tm->weight / tm->k;
	if (tm->model == TET_LINEAR) { jacobi_svd(3, 3, F, U, S, V); return 0.5 * tm->k * vol * ((S[0] - 1) * (S[0] - 1) + (S[1] - 1) * (S[1] - 1) + (S[2] - 1) * (S[2] - 1)); }
	{ prox_problem p; memset(&p, 0, sizeof(p)); p.model = tm->model; p.mu = tm->mu; p.lambda = tm->lambda; p.kappa = tm->kappa; p.k = tm->k; signed_svd(F, S, U, V); memcpy(p.x0, S, sizeof(S)); if (S[2] < 0) S[2] = -S[2]; return prox_value(&p, S) * vol; }
}

/* stand-alone prox on n column-major deformation gradients */
int oracle_prox_tets(int model, double mu, double lambda, double kappa, int n, const double *z_in, double *z_out)
{
	int i, bad = 0;
	#pragma omp parallel for reduction(+:bad)
	for (i = 0; i < n; ++i) {
		double z[9]; memcpy(z, z_in + 9 * i, sizeof(z));
		if (model == TET_LINEAR) prox_tet_linear(z);
		else { prox_problem p; memset(&p, 0, sizeof(p)); p.model = model; p.mu = mu; p.lambda = lambda; p.kappa = kappa; p.k = lambda + (2.0 / 3.0) * mu; bad += prox_tet_hyper(&p, z); }
		memcpy(z_out + 9 * i, z, sizeof(z));
	}
	return bad;
}
/* the same with the prox penalty K given separately (a SplineTet whose spline constants differ from its Lame) */
int oracle_prox_tets_k(int model, double mu, double lambda, double kappa, double K, int n, const double *z_in, double *z_out)
{
	int i, bad = 0;
	#pragma omp parallel for reduction(+:bad)
	for (i = 0; i < n; ++i) {
		double z[9]; memcpy(z, z_in + 9 * i, sizeof(z));
		if (model == TET_LINEAR) prox_tet_linear(z);
		else { prox_problem p; memset(&p, 0, sizeof(p)); p.model = model; p.mu = mu; p.lambda = lambda; p.kappa = kappa; p.k = K; bad += prox_tet_hyper(&p, z); }
		memcpy(z_out + 9 * i, z, sizeof(z));
	}
	return bad;
}
int oracle_prox_tris(double limit_min, double limit_max, int n, const double *z_in, double *z_out)
{
	int i;
	for (i = 0; i < n; ++i) { double z[6]; memcpy(z, z_in + 6 * i, sizeof(z)); prox_tri(limit_min, limit_max, z); memcpy(z_out + 6 * i, z, sizeof(z)); }
	return 0;
}
int oracle_svd3(const double *F, double *S, double *U, double *V) { signed_svd(F, S, U, V); return 0; }
int oracle_omp_threads(void) { return omp_get_max_threads(); }
