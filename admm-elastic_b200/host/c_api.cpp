// c_api.cpp -- extern "C" view of the C++ host mirror (admm_b200::Solver) so that Python (ctypes),
// bench.py and the parity tests drive the same host code a C++ application would.
// Error convention: 0 ok, 1 exception (message from admmhost_last_error), 2 initialize()==false.
#include "admm_b200.hpp"
#include <cstring>

using namespace admm_b200;

namespace {
struct Host {
	Solver solver;
	std::string error;
};
template <typename F> int guarded(Host *h, F f) {
	try { f(); return 0; }
	catch (std::exception &e) { h->error = e.what(); return 1; }
}
Lame make_lame(double mu, double lambda, double lmin, double lmax) { Lame l; l.mu = mu; l.lambda = lambda; l.limit_min = lmin; l.limit_max = lmax; return l; }
}

extern "C" {

void *admmhost_create() { return new Host(); }
void admmhost_destroy(void *h) { delete (Host *)h; }
const char *admmhost_last_error(void *h) { return ((Host *)h)->error.c_str(); }

int admmhost_add_nodes(void *h_, const double *x, const double *m, int n_verts) {
	Host *h = (Host *)h_;
	return h->solver.add_nodes(const_cast<double *>(x), const_cast<double *>(m), n_verts);
}

// create_tets_from_mesh<double, TYPE> (src/TetEnergyTerm.hpp:35-51); model = admm_b200_tet_model
int admmhost_add_tets(void *h_, const double *verts, const int *inds, int n_tets, int model, double mu, double lambda, double kappa, int vertex_offset) {
	Host *h = (Host *)h_;
	return guarded(h, [&]() {
		Lame lame = make_lame(mu, lambda, -100.0, 100.0);
		auto &et = h->solver.energyterms;
		switch (model) {
		case ADMM_B200_TET_LINEAR: create_tets_from_mesh<double, TetEnergyTerm>(et, verts, inds, n_tets, lame, vertex_offset); break;
		case ADMM_B200_TET_NEOHOOKEAN: create_tets_from_mesh<double, NeoHookeanTet>(et, verts, inds, n_tets, lame, vertex_offset); break;
		case ADMM_B200_TET_STVK: create_tets_from_mesh<double, StVKTet>(et, verts, inds, n_tets, lame, vertex_offset); break;
		case ADMM_B200_TET_SPLINE_NH: case ADMM_B200_TET_SPLINE_STVK: case ADMM_B200_TET_SPLINE_COROT: {
			std::shared_ptr<xu::Spline> sp = std::make_shared<xu::Spline>((xu::Spline::Type)(model - ADMM_B200_TET_SPLINE_NH), mu, lambda, kappa);
			for (int i = 0; i < n_tets; ++i) {
				Vec4i tet(inds[i * 4], inds[i * 4 + 1], inds[i * 4 + 2], inds[i * 4 + 3]);
				std::vector<Vec3> tv;
				for (int c = 0; c < 4; ++c) tv.emplace_back(verts[tet[c] * 3], verts[tet[c] * 3 + 1], verts[tet[c] * 3 + 2]);
				for (int c = 0; c < 4; ++c) tet[c] += vertex_offset;
				et.emplace_back(std::make_shared<SplineTet>(tet, tv, lame, sp));
			}
		} break;
		default: throw std::runtime_error("unknown tet model");
		}
	});
}

// SplineTet with a spline whose constants differ from the element's Lame (src/TetEnergyTerm.hpp:200-205)
int admmhost_add_spline_tets(void *h_, const double *verts, const int *inds, int n_tets, int spline_type, double mu, double lambda, double sp_mu, double sp_lambda, double sp_kappa, int vertex_offset) {
	Host *h = (Host *)h_;
	return guarded(h, [&]() {
		Lame lame = make_lame(mu, lambda, -100.0, 100.0);
		std::shared_ptr<xu::Spline> sp = std::make_shared<xu::Spline>((xu::Spline::Type)spline_type, sp_mu, sp_lambda, sp_kappa);
		for (int i = 0; i < n_tets; ++i) {
			Vec4i tet(inds[i * 4], inds[i * 4 + 1], inds[i * 4 + 2], inds[i * 4 + 3]);
			std::vector<Vec3> tv;
			for (int c = 0; c < 4; ++c) tv.emplace_back(verts[tet[c] * 3], verts[tet[c] * 3 + 1], verts[tet[c] * 3 + 2]);
			for (int c = 0; c < 4; ++c) tet[c] += vertex_offset;
			h->solver.energyterms.emplace_back(std::make_shared<SplineTet>(tet, tv, lame, sp));
		}
	});
}

int admmhost_add_tris(void *h_, const double *verts, const int *inds, int n_tris, double mu, double lambda, double limit_min, double limit_max, int vertex_offset) {
	Host *h = (Host *)h_;
	return guarded(h, [&]() {
		Lame lame = make_lame(mu, lambda, limit_min, limit_max);
		create_tris_from_mesh<double, TriEnergyTerm>(h->solver.energyterms, verts, inds, n_tris, lame, vertex_offset);
	});
}

int admmhost_set_pins(void *h_, const int *inds, const double *points, int n) {
	Host *h = (Host *)h_;
	return guarded(h, [&]() {
		std::vector<int> i(inds, inds + n);
		std::vector<Vec3> p;
		if (points) for (int k = 0; k < n; ++k) p.emplace_back(points[3 * k], points[3 * k + 1], points[3 * k + 2]);
		h->solver.set_pins(i, p);
	});
}

int admmhost_add_floor(void *h_, double y) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.add_obstacle(std::make_shared<Floor>(y)); }); }
int admmhost_add_sphere(void *h_, const double *c, double r) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.add_obstacle(std::make_shared<Sphere>(Vec3(c[0], c[1], c[2]), r)); }); }

int admmhost_set_options(void *h_, int device, int precision, int gs_max_iters, double gs_tol, double gs_omega, int coloring, int keep_z, int timers, void *stream) {
	Host *h = (Host *)h_;
	Solver::DeviceOptions &o = h->solver.device_options;
	o.device = device; o.precision = precision; o.gs_max_iters = gs_max_iters; o.gs_tol = gs_tol; o.gs_omega = gs_omega;
	o.coloring = coloring; o.keep_z = keep_z != 0; o.timers = timers != 0; o.stream = stream;
	return 0;
}

int admmhost_set_surface_inds(void *h_, const int *idx, int n) { ((Host *)h_)->solver.surface_inds.assign(idx, idx + n); return 0; }
// Solver::ext_forces.push_back(std::make_shared<WindForce>(tris)); returns its index in ext_forces
int admmhost_add_wind(void *h_, const int *tris, int n_tris, const double *dir) {
	Host *h = (Host *)h_;
	std::shared_ptr<admm_b200::WindForce> w = std::make_shared<admm_b200::WindForce>(std::vector<int>(tris, tris + 3 * (size_t)n_tris));
	w->direction = admm_b200::Vec3(dir[0], dir[1], dir[2]);
	h->solver.ext_forces.emplace_back(w);
	return (int)h->solver.ext_forces.size() - 1;
}
int admmhost_set_wind_direction(void *h_, int index, const double *dir) {
	Host *h = (Host *)h_;
	if (index < 0 || index >= (int)h->solver.ext_forces.size()) return 1;
	admm_b200::WindForce *w = dynamic_cast<admm_b200::WindForce *>(h->solver.ext_forces[index].get());
	if (!w) return 1;
	w->direction = admm_b200::Vec3(dir[0], dir[1], dir[2]);
	return 0;
}
// WindForce::project alone (host form of the device kernels)
void admmhost_wind_project(const int *tris, int n_tris, const double *dir, double dt, int n_nodes, const double *x, double *v) {
	admm_b200::WindForce w(std::vector<int>(tris, tris + 3 * (size_t)n_tris));
	w.direction = admm_b200::Vec3(dir[0], dir[1], dir[2]);
	std::vector<double> xx(x, x + 3 * (size_t)n_nodes), vv(v, v + 3 * (size_t)n_nodes), mm(3 * (size_t)n_nodes, 1.0);
	w.project(dt, xx, vv, mm);
	for (size_t i = 0; i < vv.size(); ++i) v[i] = vv[i];
}
int admmhost_set_gs_parts(void *h_, int n_parts) { ((Host *)h_)->solver.device_options.gs_parts = n_parts; return 0; }

int admmhost_set_rank(void *h_, int rank, int world) {
	Host *h = (Host *)h_;
	h->solver.device_options.rank = rank; h->solver.device_options.world = world;
	return 0;
}
int admmhost_mgpu_export(void *h_, void *blob) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.mgpu_export(blob); }); }
int admmhost_mgpu_import(void *h_, int peer, const void *blob) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.mgpu_import(peer, blob); }); }
int admmhost_mgpu_ready(void *h_) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.mgpu_ready(); }); }
int admmhost_mgpu_nodes(void *h_, int *out2) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.mgpu_nodes(out2[0], out2[1]); }); }
int admmhost_get_node_owner(void *h_, int *out) {
	const std::vector<int> &o = ((Host *)h_)->solver.node_owner();
	for (size_t i = 0; i < o.size(); ++i) out[i] = o[i];
	return (int)o.size();
}

int admmhost_set_colors(void *h_, int n_colors, const int *offsets, const int *nodes) {
	Host *h = (Host *)h_;
	h->solver.user_colors.clear();
	for (int c = 0; c < n_colors; ++c) h->solver.user_colors.emplace_back(nodes + offsets[c], nodes + offsets[c + 1]);
	h->solver.device_options.coloring = 2;
	return 0;
}

int admmhost_initialize(void *h_, double dt, int admm_iters, double gravity, int linsolver, double constraint_w) {
	Host *h = (Host *)h_;
	int rc = 0;
	int e = guarded(h, [&]() {
		Solver::Settings s;
		s.timestep_s = dt; s.verbose = 0; s.admm_iters = admm_iters; s.gravity = gravity; s.linsolver = linsolver; s.constraint_w = constraint_w;
		if (!h->solver.initialize(s)) rc = 2;
	});
	return e ? e : rc;
}

int admmhost_set_admm_iters(void *h_, int it) {
	// Settings are copied at initialize; tests re-run with other iteration counts without a rebuild.
	Host *h = (Host *)h_;
	const_cast<Solver::Settings &>(h->solver.settings()).admm_iters = it;
	return 0;
}

int admmhost_step(void *h_) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.step(); }); }
int admmhost_step_device(void *h_) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.step_device(); }); }
int admmhost_upload_state(void *h_) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.upload_state(); }); }
int admmhost_sync_state(void *h_) { Host *h = (Host *)h_; return guarded(h, [&]() { h->solver.sync_state(); }); }

void admmhost_runtime(void *h_, double *out) {
	const Solver::RuntimeData &r = ((Host *)h_)->solver.runtime_data();
	out[0] = r.global_ms; out[1] = r.local_ms; out[2] = r.collision_ms; out[3] = r.inner_iters; out[4] = r.assemble_ms; out[5] = r.step_ms;
}
int admmhost_dof(void *h_) { return (int)((Host *)h_)->solver.m_x.size(); }
int admmhost_n_terms(void *h_) { return (int)((Host *)h_)->solver.energyterms.size(); }
int admmhost_n_rows(void *h_) { return ((Host *)h_)->solver.n_reduction_rows(); }
void admmhost_get_x(void *h_, double *x) { auto &s = ((Host *)h_)->solver; std::memcpy(x, s.m_x.data(), sizeof(double) * s.m_x.size()); }
void admmhost_get_v(void *h_, double *v) { auto &s = ((Host *)h_)->solver; std::memcpy(v, s.m_v.data(), sizeof(double) * s.m_v.size()); }
void admmhost_set_x(void *h_, const double *x) { auto &s = ((Host *)h_)->solver; std::memcpy(s.m_x.data(), x, sizeof(double) * s.m_x.size()); }
void admmhost_set_v(void *h_, const double *v) { auto &s = ((Host *)h_)->solver; std::memcpy(s.m_v.data(), v, sizeof(double) * s.m_v.size()); }
double *admmhost_x_ptr(void *h_) { return ((Host *)h_)->solver.m_x.data(); }

// scalar system matrix L (A = L (x) I3 + M) and the colours chosen at initialize
void admmhost_system_shape(void *h_, long long *out) { const sparse::Csr &A = ((Host *)h_)->solver.system_matrix(); out[0] = A.n; out[1] = (long long)A.cols.size(); }
void admmhost_system_get(void *h_, int *rowptr, int *cols, double *vals) {
	const sparse::Csr &A = ((Host *)h_)->solver.system_matrix();
	std::memcpy(rowptr, A.rowptr.data(), sizeof(int) * A.rowptr.size());
	std::memcpy(cols, A.cols.data(), sizeof(int) * A.cols.size());
	std::memcpy(vals, A.vals.data(), sizeof(double) * A.vals.size());
}
int admmhost_n_colors(void *h_) { return (int)((Host *)h_)->solver.colors().size(); }
void admmhost_get_colors(void *h_, int *offsets, int *nodes) {
	const auto &c = ((Host *)h_)->solver.colors();
	int k = 0; offsets[0] = 0;
	for (size_t i = 0; i < c.size(); ++i) { for (int v : c[i]) nodes[k++] = v; offsets[i + 1] = k; }
}
// g_index of every energy term, in energyterms order (after initialize)
void admmhost_get_row_offsets(void *h_, int *out) { auto &et = ((Host *)h_)->solver.energyterms; for (size_t i = 0; i < et.size(); ++i) out[i] = et[i]->global_index(); }
// Rest data of the tet / triangle terms exactly as handed to the device, in energyterms order (parity tests compare them
// with what the reference-side binding harvests from the reference's get_reduction triplets)
int admmhost_get_tet_rest(void *h_, int *idx4, double *dminv9, double *w, int *row) {
	int e = 0;
	for (auto &t : ((Host *)h_)->solver.energyterms) {
		if (t->kind() != EnergyTerm::TET) continue;
		TetEnergyTerm *tt = static_cast<TetEnergyTerm *>(t.get());
		if (idx4) { for (int c = 0; c < 4; ++c) idx4[4 * e + c] = tt->indices()[c]; for (int k = 0; k < 9; ++k) dminv9[9 * e + k] = tt->rest_inverse()[k]; w[e] = tt->get_weight(); row[e] = tt->global_index(); }
		++e;
	}
	return e;
}
int admmhost_get_tri_rest(void *h_, int *idx3, double *rest4, double *w, int *row) {
	int e = 0;
	for (auto &t : ((Host *)h_)->solver.energyterms) {
		if (t->kind() != EnergyTerm::TRI) continue;
		TriEnergyTerm *tt = static_cast<TriEnergyTerm *>(t.get());
		if (idx3) { for (int c = 0; c < 3; ++c) idx3[3 * e + c] = tt->indices()[c]; for (int k = 0; k < 4; ++k) rest4[4 * e + k] = tt->rest_inverse()[k]; w[e] = tt->get_weight(); row[e] = tt->global_index(); }
		++e;
	}
	return e;
}
void *admmhost_device_handle(void *h_) { return ((Host *)h_)->solver.device_handle(); }

// Host-only helpers exposed for CPU tests (no device needed)
int admmhost_color_matrix(int n, const int *rowptr, const int *cols, const double *vals, int method, int *n_colors_out, int *offsets, int *nodes) {
	try {
		sparse::Csr A; A.n = n; A.rowptr.assign(rowptr, rowptr + n + 1); A.cols.assign(cols, cols + rowptr[n]); A.vals.assign(vals, vals + rowptr[n]);
		std::vector<std::vector<int>> colors;
		if (method == 1) sparse::color_random_palette(A, colors); else sparse::color_greedy(A, colors);
		if (!sparse::coloring_is_valid(A, colors)) return 3;
		*n_colors_out = (int)colors.size();
		int k = 0; offsets[0] = 0;
		for (size_t i = 0; i < colors.size(); ++i) { for (int v : colors[i]) nodes[k++] = v; offsets[i + 1] = k; }
		return 0;
	} catch (std::exception &) { return 1; }
}

// L D L^T of (A) with nested-dissection ordering; returns the solution of A x = b computed on the
// host from the factor (forward/diagonal/backward), used by CPU tests of the factorisation only.
int admmhost_ldlt_check(int n, const int *rowptr, const int *cols, const double *vals, const double *pos3, const double *b, double *x, long long *stats) {
	try {
		sparse::Csr A; A.n = n; A.rowptr.assign(rowptr, rowptr + n + 1); A.cols.assign(cols, cols + rowptr[n]); A.vals.assign(vals, vals + rowptr[n]);
		std::vector<int> perm = sparse::order_nested_dissection(A, pos3);
		sparse::Ldlt f = sparse::factor_ldlt(A, perm);
		std::vector<double> y(n);
		for (int k = 0; k < n; ++k) y[k] = b[perm[k]];
		for (int j = 0; j < n; ++j) for (int p = f.Lp[j]; p < f.Lp[j + 1]; ++p) y[f.Li[p]] -= f.Lx[p] * y[j];
		for (int k = 0; k < n; ++k) y[k] /= f.D[k];
		for (int j = n - 1; j >= 0; --j) for (int p = f.Lp[j]; p < f.Lp[j + 1]; ++p) y[j] -= f.Lx[p] * y[f.Li[p]];
		for (int k = 0; k < n; ++k) x[perm[k]] = y[k];
		if (stats) { stats[0] = f.Lp[n]; }
		return 0;
	} catch (std::exception &) { return 1; }
}

// The same factor walked by the device's block plan (csrc/ldlt_blocks.hpp) on the host; stats as admm_b200_ldlt_blocks_check
int admmhost_ldlt_blocks_check(int n, const int *rowptr, const int *cols, const double *vals, const double *pos3, const double *b, double *x, long long *stats) {
	try {
		sparse::Csr A; A.n = n; A.rowptr.assign(rowptr, rowptr + n + 1); A.cols.assign(cols, cols + rowptr[n]); A.vals.assign(vals, vals + rowptr[n]);
		std::vector<int> perm = sparse::order_nested_dissection(A, pos3);
		sparse::Ldlt f = sparse::factor_ldlt(A, perm);
		return admm_b200_ldlt_blocks_check(n, f.perm.data(), f.Lp.data(), f.Li.data(), f.Lx.data(), f.D.data(), b, x, stats);
	} catch (std::exception &) { return 1; }
}

} // extern "C"
