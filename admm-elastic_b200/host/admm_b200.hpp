// admm_b200.hpp -- C++ host side of the B200-native ADMM-elastic step.
//
// Mirrors the plugin surface of the reference (mattoverby/admm-elastic @ c6c09a3) for the hot path:
//   admm::Lame, admm::EnergyTerm            src/EnergyTerm.hpp:34-107
//   admm::TetEnergyTerm / NeoHookeanTet / StVKTet / SplineTet   src/TetEnergyTerm.hpp:35-208
//   admm::TriEnergyTerm                      src/TriEnergyTerm.hpp:31-76
//   admm::SpringPin                          src/SpringEnergyTerm.hpp:31-73
//   admm::Floor / Sphere                     src/PassiveObject.hpp:32-64
//   admm::Solver                             src/Solver.hpp:33-141, src/Solver.cpp:35-261
// with the same names, argument meaning and error behaviour (std::runtime_error with the
// reference's messages; initialize() returns false + stderr on bad node data), so that code written
// against the reference reads the same here.  The reference's Eigen types are replaced by
// std::vector<double> / small PODs (Eigen is not a dependency of this build).
//
// What differs, on purpose: the terms are DESCRIPTORS.  They compute their rest-state data and
// reduction triplets on the host exactly as the reference constructors do, but prox()/update() run
// only on the GPU through the C-ABI (include/admm_b200.h).  There is no CPU implementation of the
// hot path in this library: Solver::initialize throws if no CUDA device is usable.
#pragma once
#include "../../include/admm_b200.h"
#include "sparse.hpp"
#include <cstdio>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace admm_b200 {

struct Vec3 {
	double v[3];
	Vec3() : v{0, 0, 0} {}
	Vec3(double x, double y, double z) : v{x, y, z} {}
	double &operator[](int i) { return v[i]; }
	double operator[](int i) const { return v[i]; }
	Vec3 operator-(const Vec3 &o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
	Vec3 operator+(const Vec3 &o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
	Vec3 operator*(double s) const { return Vec3(v[0] * s, v[1] * s, v[2] * s); }
	double dot(const Vec3 &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
	double norm() const { return std::sqrt(dot(*this)); }
	Vec3 normalized() const { double n = norm(); return Vec3(v[0] / n, v[1] / n, v[2] / n); }
};
struct Vec4i { int v[4]; Vec4i() : v{0, 0, 0, 0} {} Vec4i(int a, int b, int c, int d) : v{a, b, c, d} {} int &operator[](int i) { return v[i]; } int operator[](int i) const { return v[i]; } };
struct Vec3i { int v[3]; Vec3i() : v{0, 0, 0} {} Vec3i(int a, int b, int c) : v{a, b, c} {} int &operator[](int i) { return v[i]; } int operator[](int i) const { return v[i]; } };

// Eigen::Triplet<double> stand-in
struct Triplet {
	int m_row, m_col; double m_value;
	Triplet(int r, int c, double v) : m_row(r), m_col(c), m_value(v) {}
	int row() const { return m_row; } int col() const { return m_col; } double value() const { return m_value; }
};

//
//	Lame constants (src/EnergyTerm.hpp:34-59)
//
class Lame {
public:
	static Lame rubber() { return Lame(10000000, 0.499); }
	static Lame soft_rubber() { return Lame(10000000, 0.399); }
	static Lame very_soft_rubber() { return Lame(1000000, 0.299); }
	double mu, lambda;
	double bulk_modulus() const { return lambda + (2.0 / 3.0) * mu; }
	double limit_min, limit_max;
	Lame(double k, double v) : mu(k / (2.0 * (1.0 + v))), lambda(k * v / ((1.0 + v) * (1.0 - 2.0 * v))), limit_min(-100.0), limit_max(100.0) {}
	Lame() : mu(0), lambda(0), limit_min(-100.0), limit_max(100.0) {}
};

//
//	Energy term base (src/EnergyTerm.hpp:65-128)
//
class EnergyTerm {
private:
	int g_index = 0; // starting row of the reduction matrix
public:
	enum Kind { TET, TRI, PIN };
	virtual ~EnergyTerm() {}
	// Called by the solver to create the global reduction and weight matrices (src/EnergyTerm.hpp:113-128)
	inline void get_reduction(std::vector<Triplet> &triplets, std::vector<double> &weights) {
		std::vector<Triplet> temp;
		get_reduction(temp);
		g_index = (int)weights.size();
		for (const Triplet &t : temp) triplets.emplace_back(t.row() + g_index, t.col(), t.value());
		double w = get_weight();
		if (w <= 0.0) throw std::runtime_error("**EnergyTerm::get_reduction Error: Some weight leq 0");
		for (int i = 0; i < get_dim(); ++i) weights.emplace_back(w);
	}
	virtual int get_dim() const = 0;
	virtual double get_weight() const = 0;
	virtual Kind kind() const = 0;
	int global_index() const { return g_index; }
	void set_global_index(int g) { g_index = g; }
protected:
	virtual void get_reduction(std::vector<Triplet> &triplets) = 0;
};

namespace xu {
// Built-in principal-stretch splines (src/XuSpline.hpp:48-94).  Custom virtual splines cannot run
// on the GPU; only these three are accepted by SplineTet.
struct Spline {
	enum Type { NEOHOOKEAN = 0, STVK = 1, COROTATED = 2 };
	Type type; double mu, lambda, kappa;
	Spline(Type t, double mu_, double lambda_, double kappa_) : type(t), mu(mu_), lambda(lambda_), kappa(kappa_) {}
};
struct NeoHookean : Spline { NeoHookean(double mu_, double lambda_, double kappa_) : Spline(NEOHOOKEAN, mu_, lambda_, kappa_) {} };
struct StVK : Spline { StVK(double mu_, double lambda_, double kappa_) : Spline(STVK, mu_, lambda_, kappa_) {} };
struct CoRotated : Spline { CoRotated(double mu_, double lambda_, double kappa_) : Spline(COROTATED, mu_, lambda_, kappa_) {} };
}

//
//	Tet energy terms (src/TetEnergyTerm.hpp:57-208, src/TetEnergyTerm.cpp:31-71)
//
class TetEnergyTerm : public EnergyTerm {
protected:
	Vec4i tet;
	Lame lame;
	double volume, weight;
	double edges_inv[9]; // row-major: edges_inv[3c+r] = edges_inv(c,r)
public:
	int get_dim() const { return 9; }
	double get_weight() const { return weight; }
	Kind kind() const { return TET; }
	virtual int model() const { return ADMM_B200_TET_LINEAR; }
	virtual double kappa() const { return 0.0; }
	virtual double model_mu() const { return lame.mu; }
	virtual double model_lambda() const { return lame.lambda; }
	// K of the prox penalty: always the ELEMENT's Lame (problem.k = lame_.bulk_modulus(), src/TetEnergyTerm.hpp:193-200)
	double prox_bulk_modulus() const { return lame.bulk_modulus(); }
	const Vec4i &indices() const { return tet; }
	const Lame &material() const { return lame; }
	const double *rest_inverse() const { return edges_inv; }
	double rest_volume() const { return volume; }

	TetEnergyTerm(const Vec4i &tet_, const std::vector<Vec3> &verts, const Lame &lame_) : tet(tet_), lame(lame_), volume(0.0), weight(0.0) {
		// edges = [v1-v0, v2-v0, v3-v0] as columns; edges_inv = edges^-1 (src/TetEnergyTerm.cpp:35-39)
		double e[9]; // e[3r+c] = edges(r,c)
		for (int c = 0; c < 3; ++c) { Vec3 d = verts[c + 1] - verts[0]; for (int r = 0; r < 3; ++r) e[3 * r + c] = d[r]; }
		double det = e[0] * (e[4] * e[8] - e[5] * e[7]) - e[1] * (e[3] * e[8] - e[5] * e[6]) + e[2] * (e[3] * e[7] - e[4] * e[6]);
		double id = 1.0 / det;
		// inverse(r,c) by cofactors; stored as edges_inv[3r+c]
		edges_inv[0] = (e[4] * e[8] - e[5] * e[7]) * id; edges_inv[1] = (e[2] * e[7] - e[1] * e[8]) * id; edges_inv[2] = (e[1] * e[5] - e[2] * e[4]) * id;
		edges_inv[3] = (e[5] * e[6] - e[3] * e[8]) * id; edges_inv[4] = (e[0] * e[8] - e[2] * e[6]) * id; edges_inv[5] = (e[2] * e[3] - e[0] * e[5]) * id;
		edges_inv[6] = (e[3] * e[7] - e[4] * e[6]) * id; edges_inv[7] = (e[1] * e[6] - e[0] * e[7]) * id; edges_inv[8] = (e[0] * e[4] - e[1] * e[3]) * id;
		volume = det / 6.0f;
		if (volume < 0) throw std::runtime_error("**TetEnergyTerm Error: Inverted initial tet");
		double k = lame.bulk_modulus();
		weight = std::sqrt(k * volume);
	}

	void get_reduction(std::vector<Triplet> &triplets) {
		// D = S * edges_inv (4x3), rows of the term = 3r+j, value Dt(r,c) (src/TetEnergyTerm.cpp:50-71)
		double D[4][3];
		for (int r = 0; r < 3; ++r) {
			D[0][r] = -(edges_inv[r] + edges_inv[3 + r] + edges_inv[6 + r]);
			for (int c = 0; c < 3; ++c) D[c + 1][r] = edges_inv[3 * c + r];
		}
		const int rows[3] = {0, 3, 6};
		const int cols[4] = {3 * tet[0], 3 * tet[1], 3 * tet[2], 3 * tet[3]};
		for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) for (int j = 0; j < 3; ++j) triplets.emplace_back(rows[r] + j, cols[c] + j, D[c][r]);
	}
};

class HyperElasticTet : public TetEnergyTerm {
public:
	HyperElasticTet(const Vec4i &tet_, const std::vector<Vec3> &verts, const Lame &lame_) : TetEnergyTerm(tet_, verts, lame_) {}
};
class NeoHookeanTet : public HyperElasticTet {
public:
	NeoHookeanTet(const Vec4i &tet_, const std::vector<Vec3> &verts, const Lame &lame_) : HyperElasticTet(tet_, verts, lame_) {}
	int model() const { return ADMM_B200_TET_NEOHOOKEAN; }
};
class StVKTet : public HyperElasticTet {
public:
	StVKTet(const Vec4i &tet_, const std::vector<Vec3> &verts, const Lame &lame_) : HyperElasticTet(tet_, verts, lame_) {}
	int model() const { return ADMM_B200_TET_STVK; }
};
class SplineTet : public HyperElasticTet {
	std::shared_ptr<xu::Spline> spline;
public:
	// Defaults to NeoHookean if this constructor is used (src/TetEnergyTerm.hpp:194-198)
	SplineTet(const Vec4i &tet_, const std::vector<Vec3> &verts, const Lame &lame_) : HyperElasticTet(tet_, verts, lame_), spline(std::make_shared<xu::NeoHookean>(lame_.mu, lame_.lambda, 0.0)) {}
	SplineTet(const Vec4i &tet_, const std::vector<Vec3> &verts, const Lame &lame_, std::shared_ptr<xu::Spline> spline_) : HyperElasticTet(tet_, verts, lame_), spline(spline_) {
		if (!spline) throw std::runtime_error("**SplineTet Error: null spline");
	}
	int model() const { return ADMM_B200_TET_SPLINE_NH + (int)spline->type; }
	double kappa() const { return spline->kappa; }
	double model_mu() const { return spline->mu; }
	double model_lambda() const { return spline->lambda; }
};

template <typename IN_SCALAR, typename TYPE>
inline void create_tets_from_mesh(std::vector<std::shared_ptr<EnergyTerm>> &energyterms, const IN_SCALAR *verts, const int *inds, int n_tets, const Lame &lame, const int vertex_offset) {
	energyterms.reserve(energyterms.size() + n_tets);
	for (int i = 0; i < n_tets; ++i) {
		Vec4i tet(inds[i * 4 + 0], inds[i * 4 + 1], inds[i * 4 + 2], inds[i * 4 + 3]);
		std::vector<Vec3> tv;
		for (int c = 0; c < 4; ++c) tv.emplace_back(verts[tet[c] * 3 + 0], verts[tet[c] * 3 + 1], verts[tet[c] * 3 + 2]);
		for (int c = 0; c < 4; ++c) tet[c] += vertex_offset;
		energyterms.emplace_back(std::make_shared<TYPE>(tet, tv, lame));
	}
}

//
//	Triangle energy term (src/TriEnergyTerm.hpp:52-76, src/TriEnergyTerm.cpp:29-70)
//
class TriEnergyTerm : public EnergyTerm {
protected:
	Vec3i tri;
	Lame lame;
	double area, weight;
	double rest_pose[4]; // row-major 2x2: rest_pose[2c+r] = rest_pose(c,r)
public:
	int get_dim() const { return 6; }
	double get_weight() const { return weight; }
	Kind kind() const { return TRI; }
	const Vec3i &indices() const { return tri; }
	const Lame &material() const { return lame; }
	const double *rest_inverse() const { return rest_pose; }

	TriEnergyTerm(const Vec3i &tri_, const std::vector<Vec3> &verts, const Lame &lame_) : tri(tri_), lame(lame_), area(0.0), weight(0.0) {
		if (lame.limit_min > 1.0) throw std::runtime_error("**TriEnergyTerm Error: Strain limit min should be -inf to 1");
		if (lame.limit_max < 1.0) throw std::runtime_error("**TriEnergyTerm Error: Strain limit max should be 1 to inf");
		Vec3 e12 = verts[1] - verts[0], e13 = verts[2] - verts[0];
		Vec3 n1 = e12.normalized();
		Vec3 n2 = (e13 - n1 * e13.dot(n1)).normalized();
		// B = basis^T * edges (2x2)
		double b00 = n1.dot(e12), b01 = n1.dot(e13), b10 = n2.dot(e12), b11 = n2.dot(e13);
		double det = b00 * b11 - b01 * b10;
		rest_pose[0] = b11 / det; rest_pose[1] = -b01 / det; rest_pose[2] = -b10 / det; rest_pose[3] = b00 / det;
		area = det / 2.0f;
		if (area < 0) throw std::runtime_error("**TriEnergyTerm Error: Inverted initial pose");
		weight = std::sqrt(lame.bulk_modulus() * area);
	}

	void get_reduction(std::vector<Triplet> &triplets) {
		double D[3][2];
		for (int r = 0; r < 2; ++r) { D[0][r] = -(rest_pose[r] + rest_pose[2 + r]); D[1][r] = rest_pose[r]; D[2][r] = rest_pose[2 + r]; }
		int cols[3] = {3 * tri[0], 3 * tri[1], 3 * tri[2]};
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
			triplets.emplace_back(i, cols[j] + i, D[j][0]);
			triplets.emplace_back(3 + i, cols[j] + i, D[j][1]);
		}
	}
};

template <typename IN_SCALAR, typename TYPE>
inline void create_tris_from_mesh(std::vector<std::shared_ptr<EnergyTerm>> &energyterms, const IN_SCALAR *verts, const int *inds, int n_tris, const Lame &lame, const int vertex_offset) {
	energyterms.reserve(energyterms.size() + n_tris);
	for (int i = 0; i < n_tris; ++i) {
		Vec3i tri(inds[i * 3 + 0], inds[i * 3 + 1], inds[i * 3 + 2]);
		std::vector<Vec3> tv;
		for (int c = 0; c < 3; ++c) tv.emplace_back(verts[tri[c] * 3 + 0], verts[tri[c] * 3 + 1], verts[tri[c] * 3 + 2]);
		for (int c = 0; c < 3; ++c) tri[c] += vertex_offset;
		energyterms.emplace_back(std::make_shared<TYPE>(tri, tv, lame));
	}
}

//
//	SpringPin (src/SpringEnergyTerm.hpp:31-73): 6 rows reserved, 3 live
//
class SpringPin : public EnergyTerm {
protected:
	int idx; Vec3 pin; bool active; double weight;
public:
	int get_dim() const { return 6; }
	double get_weight() const { return weight; }
	Kind kind() const { return PIN; }
	void set_pin(const Vec3 &p) { pin = p; }
	void set_active(bool a) { active = a; }
	int index() const { return idx; }
	const Vec3 &position() const { return pin; }
	bool is_active() const { return active; }
	SpringPin(int idx_, const Vec3 &pin_) : idx(idx_), pin(pin_), active(true) {
		Lame lame = Lame::rubber();
		weight = std::sqrt(lame.bulk_modulus() * 2.0);
	}
	void get_reduction(std::vector<Triplet> &triplets) {
		const int col = 3 * idx;
		triplets.emplace_back(0, col + 0, 1.0); triplets.emplace_back(1, col + 1, 1.0); triplets.emplace_back(2, col + 2, 1.0);
	}
};

//
//	Passive collision objects handled inside the GS sweep (src/PassiveObject.hpp:32-64)
//
class DynamicCollision { public: virtual ~DynamicCollision() {} }; // src/DynamicObject.hpp:31-48, placeholder: see Solver::add_dynamic_collider
class PassiveCollision { public: virtual ~PassiveCollision() {} virtual int kind() const = 0; virtual void params(double *p4) const = 0; };
class Floor : public PassiveCollision { public: double m_y; Floor(double y) : m_y(y) {} int kind() const { return ADMM_B200_FLOOR; } void params(double *p) const { p[0] = m_y; p[1] = p[2] = p[3] = 0; } };
class Sphere : public PassiveCollision { public: Vec3 center; double rad; Sphere(const Vec3 &c, double r) : center(c), rad(r) {} int kind() const { return ADMM_B200_SPHERE; } void params(double *p) const { p[0] = center[0]; p[1] = center[1]; p[2] = center[2]; p[3] = rad; } };

//
//	Explicit forces (src/ExplicitForce.hpp:31-48): applied at the top of step(), before gravity (src/Solver.cpp:53-57).
//	WindForce runs on the device (csrc/kernels.cuh: wind_tri_kernel / wind_node_kernel).  Any other subclass is a host
//	callback on m_x / m_v: step() calls it before the state goes up; step_device() refuses (the state is not on the host).
//
class ExplicitForce {
public:
	typedef std::vector<double> VecX;
	virtual ~ExplicitForce() {}
	virtual void project(double dt, VecX &x, VecX &v, VecX &m) const = 0;
};
class WindForce : public ExplicitForce { // src/ExplicitForce.hpp:40-48
public:
	WindForce(const std::vector<int> &tris_) : tris(tris_) { direction = Vec3(0, 0, 0); }
	// Host form of the device kernels (every kick from the velocities before the call, added node by node in triangle
	// order); Solver::step() does not call it -- the device applies the force -- it is here for callers that do.
	void project(double dt, VecX &x, VecX &v, VecX &m) const {
		(void)m;
		const VecX v0(v);
		const size_t nt = tris.size() / 3;
		for (size_t t = 0; t < nt; ++t) {
			const int id[3] = {3 * tris[3 * t], 3 * tris[3 * t + 1], 3 * tris[3 * t + 2]};
			double r[3], a[3], b[3];
			for (int c = 0; c < 3; ++c) { r[c] = (v0[id[0] + c] + v0[id[1] + c] + v0[id[2] + c]) / 3.0 - direction[c]; a[c] = x[id[1] + c] - x[id[0] + c]; b[c] = x[id[2] + c] - x[id[0] + c]; }
			double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
			const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), area = 0.5 * len;
			if (len > 0) for (int c = 0; c < 3; ++c) n[c] /= len;
			const double vn = n[0] * r[0] + n[1] * r[1] + n[2] * r[2], k = -1000.0 * area * vn * std::fabs(vn);
			for (int j = 0; j < 3; ++j) for (int c = 0; c < 3; ++c) v[id[j] + c] += k * n[c] * 0.33 * dt;
		}
	}
	std::vector<int> tris;
	Vec3 direction;
};

//
//	The main solver (src/Solver.hpp:33-141)
//
class Solver {
public:
	typedef std::vector<double> VecX;

	struct Settings { // src/Solver.hpp:39-50
		bool parse_args(int argc, char **argv);
		void help();
		double timestep_s; int verbose; int admm_iters; double gravity; int linsolver; double constraint_w;
		Settings() : timestep_s(1.0 / 24.0), verbose(1), admm_iters(10), gravity(-9.8), linsolver(0), constraint_w(-1) {}
	};
	struct RuntimeData { // src/Solver.hpp:54-61
		double global_ms, local_ms, collision_ms; int inner_iters;
		double assemble_ms, step_ms; // not in the reference: assembly share of global_ms, device time of the step
		RuntimeData() : global_ms(0), local_ms(0), collision_ms(0), inner_iters(0), assemble_ms(0), step_ms(0) {}
		void print(const Settings &settings);
	};
	// GPU-side knobs that have no counterpart in Settings (kept out of it to preserve its layout)
	struct DeviceOptions {
		int device = 0;
		int precision = ADMM_B200_FP32;  // element data: fp32 (default) or fp64 (validation)
		int gs_max_iters = 30;           // NodalMultiColorGS::max_iters
		double gs_tol = 1e-10;           // NodalMultiColorGS::m_tol (<=0: skip the residual test)
		double gs_omega = 1.9;           // NodalMultiColorGS::m_omega
		int coloring = 0;                // 0 greedy, 1 randomised palette (reference-like colour count), 2 user supplied
		bool keep_z = false;             // keep z on the device for debugging
		bool timers = true;              // fill RuntimeData from CUDA events (synchronises every step)
		void *stream = nullptr;          // cudaStream_t to run on (NULL: the solver's own)
		// multi-GPU: this process is rank `rank` of `world` (one process per GPU, <= 8).  Every rank builds the
		// SAME scene; initialize() keeps the elements that touch a node this rank owns.  After initialize()
		// the ranks swap mgpu_export() blobs and feed them to mgpu_import() (see include/admm_b200.h).
		int rank = 0, world = 1;
		int gs_parts = 0;                // parts of the resident Gauss-Seidel per GPU (0: one per SM), see admm_b200_set_gs_parts
	} device_options;

	Solver() : initialized(false), handle(nullptr) {}
	virtual ~Solver() { release_device(); }
	Solver(const Solver &) = delete;
	Solver &operator=(const Solver &) = delete;

	VecX m_x, m_v, m_masses; // per-node x3, as in the reference
	std::vector<int> surface_inds;
	std::vector<std::shared_ptr<EnergyTerm>> energyterms;
	std::vector<std::shared_ptr<ExplicitForce>> ext_forces; // src/Solver.hpp:71; fixed at initialize(), WindForce::direction may change
	std::vector<std::vector<int>> user_colors; // device_options.coloring == 2

	template <typename T> int add_nodes(T *x, T *m, int n_verts) { // src/Solver.hpp:127-141
		size_t prev_n = m_x.size(), n3 = (size_t)n_verts * 3;
		m_x.resize(prev_n + n3); m_v.resize(prev_n + n3); m_masses.resize(prev_n + n3);
		for (size_t i = 0; i < n3; ++i) { m_x[prev_n + i] = x[i]; m_v[prev_n + i] = 0.0; m_masses[prev_n + i] = m[i]; }
		return (int)((prev_n + n3) / 3);
	}
	virtual void set_pins(const std::vector<int> &inds, const std::vector<Vec3> &points = std::vector<Vec3>());
	virtual void add_obstacle(std::shared_ptr<PassiveCollision> obj) { passive_objs.emplace_back(obj); }
	// src/Solver.hpp:93: self collision (DynamicCollision, BVH) is not on the GPU path; the signature is
	// kept so that callers compile, and says so loudly instead of ignoring the collider
	virtual void add_dynamic_collider(std::shared_ptr<DynamicCollision>) { throw std::runtime_error("**admm_b200::Solver::add_dynamic_collider Error: self collision is not on the GPU path"); }
	virtual bool initialize(const Settings &settings_ = Settings());
	virtual void step();
	virtual const RuntimeData &runtime_data() { return m_runtime; }
	const Settings &settings() { return m_settings; }

	// device-resident stepping for callers that do not need m_x every frame: step_device() leaves the
	// state on the GPU, sync_state() brings m_x / m_v back.
	// step_device() does NOT look at m_x / m_v: after writing to them call upload_state() first.
	void step_device();
	void sync_state();
	void upload_state();
	admm_b200_solver *device_handle() { return handle; }
	// multi-GPU plumbing (device_options.world > 1)
	void mgpu_export(void *blob) { check(admm_b200_mgpu_export(handle, blob), "mgpu_export"); }
	void mgpu_import(int peer_rank, const void *blob) { check(admm_b200_mgpu_import(handle, peer_rank, blob), "mgpu_import"); }
	void mgpu_ready() { check(admm_b200_mgpu_ready(handle), "mgpu_ready"); }
	const std::vector<int> &node_owner() const { return m_node_owner; } // rank owning each node (empty when world == 1)
	// With several ranks step() moves only this rank's nodes (owned + ghost) up and down.  m_x / m_v entries of all other
	// nodes are left as they were: merge the ranks' arrays by node_owner().
	void mgpu_nodes(int &n_owned, int &n_ghost) { check(admm_b200_mgpu_nodes(handle, &n_owned, &n_ghost), "mgpu_nodes"); }
	const sparse::Csr &system_matrix() const { return scalarL; }
	const std::vector<std::vector<int>> &colors() const { return m_colors; }
	int n_reduction_rows() const { return n_D_rows; }

protected:
	Settings m_settings;
	RuntimeData m_runtime;
	bool initialized;
	admm_b200_solver *handle;
	std::unordered_map<int, Vec3> pins; // ConstraintSet::pins
	std::map<int, std::shared_ptr<SpringPin>> m_pin_energies;
	std::vector<int> pin_order; // order in which SpringPins were handed to the device
	std::vector<std::shared_ptr<PassiveCollision>> passive_objs;
	sparse::Csr scalarL;
	std::vector<std::vector<int>> m_colors;
	int n_D_rows = 0;
	bool state_on_device_newer = false;
	std::vector<int> m_node_owner;
	std::vector<int> m_wind_id; // per ext_forces entry: device id of a WindForce, -1 for a host force
	void apply_ext_forces(bool host_state);

	bool host_pinned = false;
	void release_device() {
		if (!handle) return;
		if (host_pinned) { admm_b200_unpin_host(handle, m_x.data()); admm_b200_unpin_host(handle, m_v.data()); host_pinned = false; }
		admm_b200_destroy(handle); handle = nullptr;
	}
	void check(int rc, const char *what) {
		if (rc) { std::stringstream ss; ss << "**admm_b200 " << what << ": " << admm_b200_last_error(handle); throw std::runtime_error(ss.str()); }
	}
	void push_gs_pins();
	void push_energy_pins();
};

// ---------------------------------------------------------------------------------------------
// Implementation
// ---------------------------------------------------------------------------------------------

inline void Solver::set_pins(const std::vector<int> &inds, const std::vector<Vec3> &points) { // src/Solver.cpp:113-157
	int n_pins = (int)inds.size();
	const int dof = (int)m_x.size();
	bool pin_in_place = (int)points.size() != n_pins;
	if ((dof == 0 && pin_in_place) || (pin_in_place && points.size() > 0)) throw std::runtime_error("**Solver::set_pins Error: Bad input.");
	if (pin_in_place && state_on_device_newer) sync_state();
	pins.clear();
	for (int i = 0; i < n_pins; ++i) {
		int idx = inds[i];
		if (pin_in_place) pins[idx] = Vec3(m_x[idx * 3], m_x[idx * 3 + 1], m_x[idx * 3 + 2]);
		else pins[idx] = points[i];
	}
	if (initialized && (m_settings.linsolver == 0 || m_settings.linsolver == 2)) {
		for (auto &kv : m_pin_energies) kv.second->set_active(false);
		for (int i = 0; i < n_pins; ++i) {
			int idx = inds[i];
			auto it = m_pin_energies.find(idx);
			if (it == m_pin_energies.end()) { std::stringstream err; err << "**Solver::set_pins Error: Constraint for " << idx << " not found.\n"; throw std::runtime_error(err.str()); }
			it->second->set_active(true);
			it->second->set_pin(pins[idx]);
		}
		push_energy_pins();
	} else if (initialized && m_settings.linsolver == 1) push_gs_pins();
}

inline void Solver::push_gs_pins() {
	std::vector<int> idx; std::vector<double> pos;
	std::vector<int> keys;
	for (auto &kv : pins) keys.push_back(kv.first);
	std::sort(keys.begin(), keys.end());
	for (int k : keys) { idx.push_back(k); const Vec3 &p = pins[k]; pos.push_back(p[0]); pos.push_back(p[1]); pos.push_back(p[2]); }
	check(admm_b200_set_gs_pins(handle, (int)idx.size(), idx.data(), pos.data()), "set_gs_pins");
}

inline void Solver::push_energy_pins() {
	std::vector<double> pos; std::vector<unsigned char> act;
	for (int idx : pin_order) { auto &p = m_pin_energies[idx]; pos.push_back(p->position()[0]); pos.push_back(p->position()[1]); pos.push_back(p->position()[2]); act.push_back(p->is_active() ? 1 : 0); }
	check(admm_b200_update_pins(handle, (int)pin_order.size(), pos.data(), act.data()), "update_pins");
}

inline bool Solver::initialize(const Settings &settings_) { // src/Solver.cpp:167-261
	m_settings = settings_;
	const int dof = (int)m_x.size();
	if (m_settings.verbose > 0) std::cout << "Solver::initialize: " << std::endl;
	if (m_settings.timestep_s <= 0.0) {
		std::cerr << "\n**Solver Error: timestep set to " << m_settings.timestep_s << "s, changing to 1/24s." << std::endl;
		m_settings.timestep_s = 1.0 / 24.0;
	}
	if (!((int)m_masses.size() == dof && dof >= 3)) { std::cerr << "\n**Solver Error: Problem with node data!" << std::endl; return false; }
	m_v.assign(dof, 0.0);
	const int n_nodes = dof / 3;

	release_device();
	if (admm_b200_create(device_options.device, &handle)) {
		std::stringstream ss; ss << "**admm_b200::Solver Error: " << admm_b200_last_error(nullptr);
		throw std::runtime_error(ss.str());
	}
	if (device_options.stream) check(admm_b200_set_stream(handle, device_options.stream), "set_stream");
	if (device_options.gs_parts > 0) check(admm_b200_set_gs_parts(handle, device_options.gs_parts), "set_gs_parts");

	// Energy-based hard constraints (src/Solver.cpp:190-196).  The reference appends them again on
	// every initialize(); here terms added by an earlier initialize() are dropped first
	// (SURVEY.md App. B "initialize re-appends pins if called twice": guarded).
	for (auto &kv : m_pin_energies) {
		auto it = std::find(energyterms.begin(), energyterms.end(), std::static_pointer_cast<EnergyTerm>(kv.second));
		if (it != energyterms.end()) energyterms.erase(it);
	}
	m_pin_energies.clear(); pin_order.clear();
	if (m_settings.linsolver == 0 || m_settings.linsolver == 2) {
		std::vector<int> keys;
		for (auto &kv : pins) keys.push_back(kv.first);
		std::sort(keys.begin(), keys.end());
		for (int k : keys) { m_pin_energies[k] = std::make_shared<SpringPin>(k, pins[k]); energyterms.emplace_back(m_pin_energies[k]); }
	}

	// Reduction rows + weights: only g_index and the weights are needed from get_reduction; the
	// triplets themselves are regenerated per element below, so the 36 x n_tets list is not kept.
	std::vector<double> weights;
	{
		std::vector<Triplet> scratch;
		for (auto &t : energyterms) { scratch.clear(); t->get_reduction(scratch, weights); }
	}
	n_D_rows = (int)weights.size();

	check(admm_b200_set_nodes(handle, n_nodes, m_x.data(), nullptr, m_masses.data()), "set_nodes");
	if (device_options.keep_z) check(admm_b200_set_debug(handle, 1), "set_debug");

	// Group elements into device batches (same model + material => one kernel launch) and build the
	// scalar system matrix L = dt^2 sum_e w_e^2 d_e^T d_e  (A = L (x) I3 + M).
	const double dt2 = m_settings.timestep_s * m_settings.timestep_s;
	std::vector<sparse::Entry> entries;
	struct TetGroup { int model; double mu, lambda, kappa, bulk; std::vector<int> idx, row; std::vector<double> dminv, w; };
	struct TriGroup { double lmin, lmax; std::vector<int> idx, row; std::vector<double> rest, w; };
	std::vector<TetGroup> tgroups; std::vector<TriGroup> rgroups;
	std::vector<int> p_idx, p_row; std::vector<double> p_pos, p_w;
	for (auto &term : energyterms) {
		switch (term->kind()) {
		case EnergyTerm::TET: {
			TetEnergyTerm *t = static_cast<TetEnergyTerm *>(term.get());
			TetGroup *g = nullptr;
			if (!tgroups.empty()) { TetGroup &b = tgroups.back(); if (b.model == t->model() && b.mu == t->model_mu() && b.lambda == t->model_lambda() && b.kappa == t->kappa() && b.bulk == t->prox_bulk_modulus()) g = &b; }
			if (!g) { tgroups.emplace_back(); g = &tgroups.back(); g->model = t->model(); g->mu = t->model_mu(); g->lambda = t->model_lambda(); g->kappa = t->kappa(); g->bulk = t->prox_bulk_modulus(); }
			const double *bi = t->rest_inverse();
			for (int c = 0; c < 4; ++c) g->idx.push_back(t->indices()[c]);
			for (int k = 0; k < 9; ++k) g->dminv.push_back(bi[k]);
			g->w.push_back(t->get_weight()); g->row.push_back(t->global_index());
			double D[4][3];
			for (int r = 0; r < 3; ++r) { D[0][r] = -(bi[r] + bi[3 + r] + bi[6 + r]); for (int c = 0; c < 3; ++c) D[c + 1][r] = bi[3 * c + r]; }
			double w2 = dt2 * t->get_weight() * t->get_weight();
			for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b)
				entries.push_back({t->indices()[a], t->indices()[b], w2 * (D[a][0] * D[b][0] + D[a][1] * D[b][1] + D[a][2] * D[b][2])});
		} break;
		case EnergyTerm::TRI: {
			TriEnergyTerm *t = static_cast<TriEnergyTerm *>(term.get());
			TriGroup *g = nullptr;
			if (!rgroups.empty()) { TriGroup &b = rgroups.back(); if (b.lmin == t->material().limit_min && b.lmax == t->material().limit_max) g = &b; }
			if (!g) { rgroups.emplace_back(); g = &rgroups.back(); g->lmin = t->material().limit_min; g->lmax = t->material().limit_max; }
			const double *rp = t->rest_inverse();
			for (int c = 0; c < 3; ++c) g->idx.push_back(t->indices()[c]);
			for (int k = 0; k < 4; ++k) g->rest.push_back(rp[k]);
			g->w.push_back(t->get_weight()); g->row.push_back(t->global_index());
			double D[3][2];
			for (int r = 0; r < 2; ++r) { D[0][r] = -(rp[r] + rp[2 + r]); D[1][r] = rp[r]; D[2][r] = rp[2 + r]; }
			double w2 = dt2 * t->get_weight() * t->get_weight();
			for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
				entries.push_back({t->indices()[a], t->indices()[b], w2 * (D[a][0] * D[b][0] + D[a][1] * D[b][1])});
		} break;
		case EnergyTerm::PIN: {
			SpringPin *p = static_cast<SpringPin *>(term.get());
			p_idx.push_back(p->index()); p_row.push_back(p->global_index()); p_w.push_back(p->get_weight());
			for (int j = 0; j < 3; ++j) p_pos.push_back(p->position()[j]);
			pin_order.push_back(p->index());
			entries.push_back({p->index(), p->index(), dt2 * p->get_weight() * p->get_weight()});
		} break;
		}
	}
	for (int i = 0; i < n_nodes; ++i) entries.push_back({i, i, 0.0}); // every node gets a diagonal entry
	scalarL = sparse::from_entries(n_nodes, entries);
	entries.clear(); entries.shrink_to_fit();

	// Multi-GPU: colour and partition the GLOBAL matrix (identically on every rank), then keep only the
	// elements that touch a node this rank owns; elements on a cut are computed by both sides.
	m_node_owner.clear();
	const int world = device_options.world, rank = device_options.rank;
	if (world > 1) {
		if (m_settings.linsolver != 1) throw std::runtime_error("**admm_b200::Solver Error: multi-GPU needs the NodalMultiColorGS solver (-ls 1)");
		if (!p_idx.empty()) throw std::runtime_error("**admm_b200::Solver Error: multi-GPU has no energy-based pins (they belong to LDLT / UzawaCG)");
		check(admm_b200_set_rank(handle, rank, world), "set_rank");
		const int sms = admm_b200_gs_parts(handle);
		std::vector<int> part(n_nodes);
		if (admm_b200_plan_parts(n_nodes, scalarL.rowptr.data(), scalarL.cols.data(), scalarL.vals.data(), m_x.data(), world * sms, part.data()))
			throw std::runtime_error(std::string("**admm_b200 plan_parts: ") + admm_b200_last_error(nullptr));
		m_node_owner.resize(n_nodes);
		for (int i = 0; i < n_nodes; ++i) m_node_owner[i] = part[i] / sms;
		for (auto &g : tgroups) {
			size_t keep = 0;
			const size_t n_e = g.w.size();
			for (size_t e = 0; e < n_e; ++e) {
				bool mine = false;
				for (int c = 0; c < 4; ++c) mine = mine || m_node_owner[g.idx[4 * e + c]] == rank;
				if (!mine) continue;
				for (int c = 0; c < 4; ++c) g.idx[4 * keep + c] = g.idx[4 * e + c];
				for (int k = 0; k < 9; ++k) g.dminv[9 * keep + k] = g.dminv[9 * e + k];
				g.w[keep] = g.w[e]; g.row[keep] = g.row[e];
				++keep;
			}
			g.idx.resize(4 * keep); g.dminv.resize(9 * keep); g.w.resize(keep); g.row.resize(keep);
		}
		for (auto &g : rgroups) {
			size_t keep = 0;
			const size_t n_e = g.w.size();
			for (size_t e = 0; e < n_e; ++e) {
				bool mine = false;
				for (int c = 0; c < 3; ++c) mine = mine || m_node_owner[g.idx[3 * e + c]] == rank;
				if (!mine) continue;
				for (int c = 0; c < 3; ++c) g.idx[3 * keep + c] = g.idx[3 * e + c];
				for (int k = 0; k < 4; ++k) g.rest[4 * keep + k] = g.rest[4 * e + k];
				g.w[keep] = g.w[e]; g.row[keep] = g.row[e];
				++keep;
			}
			g.idx.resize(3 * keep); g.rest.resize(4 * keep); g.w.resize(keep); g.row.resize(keep);
		}
	}
	for (auto &g : tgroups) if (!g.w.empty()) check(admm_b200_add_tets(handle, (int)g.w.size(), g.idx.data(), g.dminv.data(), g.w.data(), g.model, g.mu, g.lambda, g.kappa, g.bulk, g.row.data()), "add_tets");
	for (auto &g : rgroups) if (!g.w.empty()) check(admm_b200_add_tris(handle, (int)g.w.size(), g.idx.data(), g.rest.data(), g.w.data(), g.lmin, g.lmax, g.row.data()), "add_tris");
	if (!p_idx.empty()) check(admm_b200_add_pins(handle, (int)p_idx.size(), p_idx.data(), p_pos.data(), p_w.data(), p_row.data()), "add_pins");

	for (auto &o : passive_objs) { double p[4]; o->params(p); check(admm_b200_add_obstacle(handle, o->kind(), p), "add_obstacle"); }
	m_wind_id.assign(ext_forces.size(), -1);
	for (size_t i = 0; i < ext_forces.size(); ++i) {
		// the reference applies ext_forces in order; device and host forces cannot be interleaved, so a host force may not
		// come after a WindForce
		if (const WindForce *w = dynamic_cast<const WindForce *>(ext_forces[i].get())) {
			const double d[3] = {w->direction[0], w->direction[1], w->direction[2]};
			check(admm_b200_add_wind(handle, w->tris.data(), (int)(w->tris.size() / 3), d, &m_wind_id[i]), "add_wind");
		} else
			for (size_t j = 0; j < i; ++j) if (m_wind_id[j] >= 0) throw std::runtime_error("**admm_b200::Solver Error: a host ExplicitForce after a WindForce (put host forces first)");
	}
	if (m_settings.linsolver == 2) {
		// what UzawaCG's collision rows depend on: the candidate vertices and their order (src/Solver.cpp:93), and
		// constraint_w = 1 unless -ck overrides it (src/Solver.cpp:239,245)
		if (!surface_inds.empty()) check(admm_b200_set_surface_inds(handle, (int)surface_inds.size(), surface_inds.data()), "set_surface_inds");
		check(admm_b200_set_constraint_weight(handle, m_settings.constraint_w > 0.0 ? m_settings.constraint_w : 1.0), "set_constraint_weight");
	}

	// Linear solver (src/Solver.cpp:229-246)
	switch (m_settings.linsolver) {
	default: m_settings.linsolver = 0; // fallthrough: LDLT
	case 2: {
		if (m_settings.linsolver == 0 && !passive_objs.empty()) throw std::runtime_error("**Solver::add_obstacle Error: No collisions with LDLT solver");
		for (int i = 0; i < n_nodes; ++i) if (!(m_masses[3 * i] == m_masses[3 * i + 1] && m_masses[3 * i] == m_masses[3 * i + 2])) throw std::runtime_error("**admm_b200::Solver Error: LDLT needs equal x/y/z masses per node");
		sparse::Csr A = scalarL;
		for (int i = 0; i < n_nodes; ++i) for (int q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q) if (A.cols[q] == i) A.vals[q] += m_masses[3 * i];
		std::vector<int> perm = sparse::order_nested_dissection(A, m_x.data());
		sparse::Ldlt f = sparse::factor_ldlt(A, perm);
		check(admm_b200_set_ldlt(handle, f.n, f.perm.data(), f.Lp.data(), f.Li.data(), f.Lx.data(), f.D.data()), "set_ldlt");
	} break;
	case 1: {
		check(admm_b200_set_system(handle, scalarL.n, scalarL.rowptr.data(), scalarL.cols.data(), scalarL.vals.data()), "set_system");
		if (device_options.coloring == 2) { m_colors = user_colors; if (!sparse::coloring_is_valid(scalarL, m_colors)) throw std::runtime_error("**admm_b200::Solver Error: user colouring is not a valid colouring of A"); }
		else if (device_options.coloring == 1) sparse::color_random_palette(scalarL, m_colors);
		else sparse::color_greedy(scalarL, m_colors);
		std::vector<int> off(1, 0), nodes;
		for (auto &c : m_colors) { nodes.insert(nodes.end(), c.begin(), c.end()); off.push_back((int)nodes.size()); }
		check(admm_b200_set_colors(handle, (int)m_colors.size(), off.data(), nodes.data()), "set_colors");
		push_gs_pins();
	} break;
	}
	check(admm_b200_finalize(handle, m_settings.timestep_s, m_settings.linsolver, device_options.gs_max_iters, device_options.gs_omega, device_options.gs_tol, device_options.precision), "finalize");
	// m_x / m_v are read and written by every step(): page-lock them (they do not move after initialize)
	if (admm_b200_pin_host(handle, m_x.data(), sizeof(double) * m_x.size()) == 0) {
		if (admm_b200_pin_host(handle, m_v.data(), sizeof(double) * m_v.size()) == 0) host_pinned = true;
		else admm_b200_unpin_host(handle, m_x.data());
	}
	if (m_settings.verbose >= 1) printf("%d nodes, %d energy terms\n", (int)m_x.size() / 3, (int)energyterms.size());
	initialized = true;
	state_on_device_newer = false;
	return true;
}

inline void Solver::step() { // src/Solver.cpp:35-110
	if (!initialized) throw std::runtime_error("**Solver::step Error: not initialized");
	if (m_settings.verbose > 0) std::cout << "\nSimulating with dt: " << m_settings.timestep_s << "s..." << std::flush;
	m_runtime = RuntimeData();
	admm_b200_runtime rt;
	if (state_on_device_newer) sync_state();
	apply_ext_forces(true);
	check(admm_b200_step_host(handle, m_settings.admm_iters, m_settings.gravity, m_x.data(), m_v.data(), device_options.timers ? &rt : nullptr), "step");
	if (device_options.timers) { m_runtime.global_ms = rt.global_ms; m_runtime.local_ms = rt.local_ms; m_runtime.collision_ms = rt.collision_ms; m_runtime.inner_iters = rt.inner_iters; m_runtime.assemble_ms = rt.assemble_ms; m_runtime.step_ms = rt.step_ms; }
	if (m_settings.verbose > 0) m_runtime.print(m_settings);
}

inline void Solver::apply_ext_forces(bool host_state) { // src/Solver.cpp:53-54
	if (ext_forces.size() != m_wind_id.size()) throw std::runtime_error("**admm_b200::Solver Error: ext_forces changed after initialize");
	for (size_t i = 0; i < ext_forces.size(); ++i) {
		if (m_wind_id[i] >= 0) {
			const WindForce *w = static_cast<const WindForce *>(ext_forces[i].get());
			const double d[3] = {w->direction[0], w->direction[1], w->direction[2]};
			check(admm_b200_set_wind_direction(handle, m_wind_id[i], d), "set_wind_direction");
		} else {
			if (!host_state) throw std::runtime_error("**admm_b200::Solver Error: step_device() with a host ExplicitForce");
			ext_forces[i]->project(m_settings.timestep_s, m_x, m_v, m_masses);
		}
	}
}

inline void Solver::step_device() {
	if (!initialized) throw std::runtime_error("**Solver::step Error: not initialized");
	apply_ext_forces(false);
	m_runtime = RuntimeData();
	admm_b200_runtime rt;
	check(admm_b200_step(handle, m_settings.admm_iters, m_settings.gravity, device_options.timers ? &rt : nullptr), "step");
	if (device_options.timers) { m_runtime.global_ms = rt.global_ms; m_runtime.local_ms = rt.local_ms; m_runtime.collision_ms = rt.collision_ms; m_runtime.inner_iters = rt.inner_iters; m_runtime.assemble_ms = rt.assemble_ms; m_runtime.step_ms = rt.step_ms; }
	state_on_device_newer = true;
}

inline void Solver::upload_state() {
	if (!initialized) throw std::runtime_error("**Solver::upload_state Error: not initialized");
	check(admm_b200_upload_state(handle, m_x.data(), m_v.data()), "upload_state");
	state_on_device_newer = false;
}

inline void Solver::sync_state() {
	check(admm_b200_download_state(handle, m_x.data(), m_v.data()), "download_state");
	state_on_device_newer = false;
}

template <typename T> inline void myclamp(T &val, T min, T max) { if (val < min) val = min; if (val > max) val = max; }
inline bool Solver::Settings::parse_args(int argc, char **argv) { // src/Solver.cpp:273-294
	for (int i = 1; i < argc - 1; ++i) {
		std::string arg(argv[i]);
		std::stringstream val(argv[i + 1]);
		if (arg == "-help" || arg == "--help" || arg == "-h") { help(); return true; }
		else if (arg == "-dt") val >> timestep_s;
		else if (arg == "-v") val >> verbose;
		else if (arg == "-it") val >> admm_iters;
		else if (arg == "-g") val >> gravity;
		else if (arg == "-ls") val >> linsolver;
		else if (arg == "-ck") val >> constraint_w;
	}
	if (argc > 0) { std::string arg(argv[argc - 1]); if (arg == "-help" || arg == "--help" || arg == "-h") { help(); return true; } }
	return false;
}
inline void Solver::Settings::help() {
	printf("\n==========================================\nArgs:\n\t-dt: time step (s)\n\t-v: verbosity (higher -> show more)\n\t-it: # admm iters\n\t-g: gravity (m/s^2)\n\t-ls: linear solver (0=LDLT, 1=NCMCGS, 2=UzawaCG) \n\t-ck: constraint weights (-1 = auto) \n==========================================\n");
}
inline void Solver::RuntimeData::print(const Settings &settings) { // src/Solver.cpp:309-319
	std::cout << "\nTotal global step: " << global_ms << "ms";
	std::cout << "\nTotal local step: " << local_ms << "ms";
	std::cout << "\nTotal collision update: " << collision_ms << "ms";
	std::cout << "\nAvg global step: " << global_ms / double(settings.admm_iters) << "ms";
	std::cout << "\nAvg local step: " << local_ms / double(settings.admm_iters) << "ms";
	std::cout << "\nAvg collision update: " << collision_ms / double(settings.admm_iters) << "ms";
	std::cout << "\nADMM Iters: " << settings.admm_iters;
	std::cout << "\nAvg Inner Iters: " << float(inner_iters) / float(settings.admm_iters);
	std::cout << std::endl;
}

} // namespace admm_b200
