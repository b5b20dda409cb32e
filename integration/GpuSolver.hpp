// GpuSolver.hpp -- the reference-side binding of the B200 ADMM-elastic step.
//
// This header is what a maintainer of mattoverby/admm-elastic (@ c6c09a3) adds next to src/Solver.hpp: a subclass of
// admm::Solver (every mutator there is virtual, src/Solver.hpp:82-101) that keeps the whole plugin surface -- public
// m_x / m_v / m_masses / energyterms, add_nodes, set_pins, add_obstacle, initialize, step, runtime_data -- and runs the
// body of Solver::step() (src/Solver.cpp:35-110) on the GPU through the C-ABI of libadmm_b200.so (include/admm_b200.h).
// It is compiled against the reference's own headers (Eigen, mcloptlib, mclscene) and NOTHING of this repository but
// that C header; host and library are both C/C++, so there is no cgo/JNI layer.
//
//   admm::GpuSolver solver;                       // instead of admm::Solver
//   binding::add_tetmesh(&solver, mesh, ...);     // samples/utils/AddMeshes.hpp works unchanged (it takes a Solver*)
//   solver.set_pins(...); solver.initialize(settings);
//   while (...) { solver.step(); draw(solver.m_x); }
//
// initialize() first runs the reference's own Solver::initialize (src/Solver.cpp:167-261: SpringPins appended, D, W, A
// assembled by Eigen, the CPU linear solver built), then HARVESTS what the device needs from the public plugin surface:
//   * every term's rest data from its get_reduction() triplets (src/TetEnergyTerm.cpp:50-71, src/TriEnergyTerm.cpp:54-70,
//     src/SpringEnergyTerm.hpp:54-59), its weight from get_weight(), its first row g_index from the size of the weight
//     list before the call (src/EnergyTerm.hpp:113-128), its constitutive model from its dynamic type;
//   * the scalar system matrix from solver_termA (every reduction couples x-x, y-y, z-z only: A = L (x) I3 + M);
//   * the colour lists from the reference's own NodalMultiColorGS (graphcolor::color_matrix, GraphColor.hpp:66-72), or
//     the factor of Eigen::SimplicialLDLT (SimplicialCholesky.h) of the n x n scalar matrix for LDLT / UzawaCG;
//   * pins and Floor / Sphere obstacles from the ConstraintSet.
// Because the CPU solver is complete as well, cpu_step() runs the reference's own Solver::step() on the same object: an
// A/B switch for parity checks (tests/test_gpu_binding.py).
//
// Not supported on the GPU path (initialize throws, nothing falls back silently): custom EnergyTerm / xu::Spline
// subclasses, PassiveMesh obstacles, dynamic (self) collision, explicit forces.
#ifndef ADMM_GPUSOLVER_HPP
#define ADMM_GPUSOLVER_HPP

#include "Solver.hpp"
#include "TetEnergyTerm.hpp"
#include "TriEnergyTerm.hpp"
#include "SpringEnergyTerm.hpp"
#include "NodalMultiColorGS.hpp"
#include "UzawaCG.hpp"
#include "PassiveObject.hpp"
#include "admm_b200.h"
#include <cstdlib>
#include <sstream>

namespace admm {

class GpuSolver : public Solver {
public:
	// GPU-side knobs without a counterpart in Settings
	int device;      // CUDA device index
	int precision;   // ADMM_B200_FP32 (production) or ADMM_B200_FP64 (validation); env ADMM_B200_PRECISION=64 selects fp64
	int gs_parts;    // parts of the resident Gauss-Seidel (0 = one per SM)
	bool keep_z;     // keep z on the device (debug_get)

	// What initialize() harvested, in energyterms order per kind (public: the parity tests compare it with the mirror's)
	struct TetBatch { int model; double mu, lambda, kappa, bulk; std::vector<int> idx, row; std::vector<double> dminv, w; };
	struct TriBatch { double limit_min, limit_max; std::vector<int> idx, row; std::vector<double> rest, w; };
	std::vector<TetBatch> tet_batches;
	std::vector<TriBatch> tri_batches;
	std::vector<int> pin_idx, pin_row;
	std::vector<double> pin_w;
	std::vector< std::vector<int> > colors; // linsolver 1: the reference's own colour lists

	GpuSolver() : device(0), precision(ADMM_B200_FP32), gs_parts(0), keep_z(false), h(NULL), pinned(false) {
		const char *e = std::getenv("ADMM_B200_PRECISION");
		if( e && std::atoi(e) == 64 ){ precision = ADMM_B200_FP64; }
	}
	virtual ~GpuSolver(){ release(); }

	virtual bool initialize( const Settings &settings_=Settings() );
	virtual void step();
	virtual void set_pins( const std::vector<int> &inds, const std::vector<Vec3> &points = std::vector<Vec3>() );
	virtual void add_dynamic_collider( std::shared_ptr<DynamicCollision> ){
		throw std::runtime_error("**GpuSolver::add_dynamic_collider Error: self collision is not on the GPU path");
	}

	// The reference's own CPU step on the same object (same D, W, A, colours, pins)
	void cpu_step(){ Solver::step(); }
	admm_b200_solver *handle(){ return h; }

protected:
	admm_b200_solver *h;
	bool pinned;
	std::vector<int> pin_order; // SpringPins in the order they were handed to the device

	// NodalMultiColorGS keeps its colour lists protected (src/NodalMultiColorGS.hpp): read them through a subclass
	struct ColorPeek : public NodalMultiColorGS {
		ColorPeek( std::shared_ptr<ConstraintSet> c ) : NodalMultiColorGS(c) {}
		const std::vector< std::vector<int> > &lists() const { return A_colors; }
	};
	// Lame of a triangle term is protected (src/TriEnergyTerm.hpp:58): only the strain limits are needed from it
	struct TriPeek : public TriEnergyTerm { const Lame &material() const { return lame; } };

	void ck( int rc, const char *what ){
		if( rc ){ std::stringstream ss; ss << "**GpuSolver " << what << ": " << admm_b200_last_error(h); throw std::runtime_error(ss.str()); }
	}
	void release(){
		if( !h ){ return; }
		if( pinned ){ admm_b200_unpin_host(h, m_x.data()); admm_b200_unpin_host(h, m_v.data()); pinned = false; }
		admm_b200_destroy(h); h = NULL;
	}
	void harvest_terms();
	void push_pins();
};

//
//	Implementation
//

// Rest data of every term from the public plugin surface (see the header comment).
inline void GpuSolver::harvest_terms(){
	typedef Eigen::Triplet<double> Trip;
	tet_batches.clear(); tri_batches.clear(); pin_idx.clear(); pin_row.clear(); pin_w.clear(); pin_order.clear();
	std::vector<Trip> trips; std::vector<double> weights;
	const int n_terms = energyterms.size();
	for( int i=0; i<n_terms; ++i ){
		EnergyTerm *t = energyterms[i].get();
		const size_t t0 = trips.size();
		const int g_index = weights.size();        // the term's first row of D (src/EnergyTerm.hpp:117)
		t->get_reduction( trips, weights );          // same order as Solver::initialize => same g_index as in m_D
		const Trip *tr = &trips[t0];
		const size_t nt = trips.size()-t0;

		if( TetEnergyTerm *tet = dynamic_cast<TetEnergyTerm*>(t) ){
			// 36 triplets, index (r*4+c)*3+j: row g+3r+j, column 3*tet[c]+j, value (S*edges_inv)(c,r)  (src/TetEnergyTerm.cpp:50-71)
			if( nt != 36 || tet->get_dim() != 9 ){ throw std::runtime_error("**GpuSolver Error: unknown tet term layout"); }
			// bulk: K of the prox penalty = problem.k, the bulk modulus of the element's own Lame (src/TetEnergyTerm.hpp:125-128,193-200)
			int model = ADMM_B200_TET_LINEAR; double mu=0, lambda=0, kappa=0, bulk=0;
			if( NeoHookeanTet *nh = dynamic_cast<NeoHookeanTet*>(t) ){ model = ADMM_B200_TET_NEOHOOKEAN; mu = nh->problem.mu; lambda = nh->problem.lambda; bulk = nh->problem.k; }
			else if( StVKTet *sv = dynamic_cast<StVKTet*>(t) ){ model = ADMM_B200_TET_STVK; mu = sv->problem.mu; lambda = sv->problem.lambda; bulk = sv->problem.k; }
			else if( SplineTet *sp = dynamic_cast<SplineTet*>(t) ){
				bulk = sp->problem.k;
				xu::Spline *s = sp->problem.spline.get();
				if( xu::NeoHookean *a = dynamic_cast<xu::NeoHookean*>(s) ){ model = ADMM_B200_TET_SPLINE_NH; mu = a->mu; lambda = a->lambda; kappa = a->kappa; }
				else if( xu::StVK *b = dynamic_cast<xu::StVK*>(s) ){ model = ADMM_B200_TET_SPLINE_STVK; mu = b->mu; lambda = b->lambda; kappa = b->kappa; }
				else if( xu::CoRotated *c = dynamic_cast<xu::CoRotated*>(s) ){ model = ADMM_B200_TET_SPLINE_COROT; mu = c->mu; lambda = c->lambda; kappa = c->kappa; }
				else { throw std::runtime_error("**GpuSolver Error: custom xu::Spline subclasses cannot run on the GPU"); }
			}
			else if( dynamic_cast<HyperElasticTet*>(t) ){ throw std::runtime_error("**GpuSolver Error: custom HyperElasticTet subclasses cannot run on the GPU"); }
			if( tet_batches.empty() || tet_batches.back().model != model || tet_batches.back().mu != mu ||
				tet_batches.back().lambda != lambda || tet_batches.back().kappa != kappa || tet_batches.back().bulk != bulk ){
				TetBatch nb; nb.model = model; nb.mu = mu; nb.lambda = lambda; nb.kappa = kappa; nb.bulk = bulk;
				tet_batches.push_back( nb );
			}
			TetBatch &b = tet_batches.back();
			for( int c=0; c<4; ++c ){ b.idx.push_back( tr[c*3].col()/3 ); }
			for( int c=1; c<4; ++c ){ for( int r=0; r<3; ++r ){ b.dminv.push_back( tr[(r*4+c)*3].value() ); } } // edges_inv(c-1,r)
			b.w.push_back( t->get_weight() ); b.row.push_back( g_index );
		}
		else if( TriEnergyTerm *tri = dynamic_cast<TriEnergyTerm*>(t) ){
			// 18 triplets, index (i*3+j)*2+r: row g+3r+i, column 3*tri[j]+i, value (S*rest_pose)(j,r)  (src/TriEnergyTerm.cpp:54-70)
			if( nt != 18 || tri->get_dim() != 6 ){ throw std::runtime_error("**GpuSolver Error: unknown triangle term layout"); }
			const Lame &lame = static_cast<TriPeek*>(tri)->material();
			if( tri_batches.empty() || tri_batches.back().limit_min != lame.limit_min || tri_batches.back().limit_max != lame.limit_max ){
				TriBatch nb; nb.limit_min = lame.limit_min; nb.limit_max = lame.limit_max;
				tri_batches.push_back( nb );
			}
			TriBatch &b = tri_batches.back();
			for( int j=0; j<3; ++j ){ b.idx.push_back( tr[j*2].col()/3 ); }
			for( int c=0; c<2; ++c ){ for( int r=0; r<2; ++r ){ b.rest.push_back( tr[(c+1)*2+r].value() ); } } // rest_pose(c,r)
			b.w.push_back( t->get_weight() ); b.row.push_back( g_index );
		}
		else if( dynamic_cast<SpringPin*>(t) ){
			if( nt != 3 ){ throw std::runtime_error("**GpuSolver Error: unknown pin term layout"); }
			pin_idx.push_back( tr[0].col()/3 ); pin_row.push_back( g_index ); pin_w.push_back( t->get_weight() );
			pin_order.push_back( tr[0].col()/3 );
		}
		else { throw std::runtime_error("**GpuSolver Error: this EnergyTerm subclass has no GPU kernel"); }
	}
	if( (int)weights.size() != m_W_diag.rows() ){ throw std::runtime_error("**GpuSolver Error: reduction rows changed since initialize"); }
}

inline bool GpuSolver::initialize( const Settings &settings_ ){
	// Terms appended by an earlier initialize() would be appended again (src/Solver.cpp:190-196): drop them first
	std::unordered_map<int, std::shared_ptr<SpringPin> >::iterator pe = m_pin_energies.begin();
	for( ; pe != m_pin_energies.end(); ++pe ){
		std::shared_ptr<EnergyTerm> as_term = pe->second;
		std::vector< std::shared_ptr<EnergyTerm> >::iterator it = std::find( energyterms.begin(), energyterms.end(), as_term );
		if( it != energyterms.end() ){ energyterms.erase(it); }
	}
	m_pin_energies.clear();
	if( m_constraints->collider->dynamic_objs.size() > 0 ){ throw std::runtime_error("**GpuSolver Error: self collision is not on the GPU path"); }

	// The reference's own initialize: validates, appends SpringPins, builds D, W, A and the CPU linear solver
	if( !Solver::initialize( settings_ ) ){ return false; }
	const int dof = m_x.rows(), n = dof/3;

	release();
	if( admm_b200_create( device, &h ) ){
		std::stringstream ss; ss << "**GpuSolver Error: " << admm_b200_last_error(NULL); // no CUDA device: throws, there is no CPU fallback
		throw std::runtime_error( ss.str() );
	}
	if( gs_parts > 0 ){ ck( admm_b200_set_gs_parts( h, gs_parts ), "set_gs_parts" ); }
	if( keep_z ){ ck( admm_b200_set_debug( h, 1 ), "set_debug" ); }
	ck( admm_b200_set_nodes( h, n, m_x.data(), NULL, m_masses.data() ), "set_nodes" );

	harvest_terms();
	for( size_t b=0; b<tet_batches.size(); ++b ){
		TetBatch &t = tet_batches[b];
		ck( admm_b200_add_tets( h, t.w.size(), t.idx.data(), t.dminv.data(), t.w.data(), t.model, t.mu, t.lambda, t.kappa, t.bulk, t.row.data() ), "add_tets" );
	}
	for( size_t b=0; b<tri_batches.size(); ++b ){
		TriBatch &t = tri_batches[b];
		ck( admm_b200_add_tris( h, t.w.size(), t.idx.data(), t.rest.data(), t.w.data(), t.limit_min, t.limit_max, t.row.data() ), "add_tris" );
	}
	if( pin_idx.size() ){
		std::vector<double> pos;
		for( size_t i=0; i<pin_idx.size(); ++i ){ const Vec3 &p = m_constraints->pins[ pin_idx[i] ]; pos.push_back(p[0]); pos.push_back(p[1]); pos.push_back(p[2]); }
		ck( admm_b200_add_pins( h, pin_idx.size(), pin_idx.data(), pos.data(), pin_w.data(), pin_row.data() ), "add_pins" );
	}

	// Passive obstacles (src/PassiveObject.hpp:32-64), in passive_objs order
	const std::vector< std::shared_ptr<PassiveCollision> > &objs = m_constraints->collider->passive_objs;
	for( size_t i=0; i<objs.size(); ++i ){
		double p[4] = {0,0,0,0};
		if( Floor *f = dynamic_cast<Floor*>(objs[i].get()) ){ p[0] = f->m_y; ck( admm_b200_add_obstacle( h, ADMM_B200_FLOOR, p ), "add_obstacle" ); }
		else if( Sphere *s = dynamic_cast<Sphere*>(objs[i].get()) ){ p[0]=s->center[0]; p[1]=s->center[1]; p[2]=s->center[2]; p[3]=s->rad; ck( admm_b200_add_obstacle( h, ADMM_B200_SPHERE, p ), "add_obstacle" ); }
		else { throw std::runtime_error("**GpuSolver Error: only Floor and Sphere obstacles are on the GPU path"); }
	}

	// The scalar matrix: rows / columns 0, 3, 6, ... of solver_termA (src/Solver.cpp:226) minus the mass diagonal
	std::vector<int> rowptr(n+1,0), cols; std::vector<double> vals;
	for( int i=0; i<n; ++i ){
		bool has_diag = false;
		for( SparseMat::InnerIterator it(solver_termA,3*i); it; ++it ){
			if( it.col()%3 != 0 ){
				if( it.value() != 0.0 ){ throw std::runtime_error("**GpuSolver Error: A is not of the form L (x) I3 (a term couples different components)"); }
				continue;
			}
			const int j = it.col()/3;
			double v = it.value();
			if( j == i ){ v -= m_masses[3*i]; has_diag = true; }
			cols.push_back(j); vals.push_back(v);
		}
		if( !has_diag ){ throw std::runtime_error("**GpuSolver Error: A has no diagonal entry for a node"); }
		rowptr[i+1] = cols.size();
	}

	if( m_settings.linsolver == 2 ){
		// what UzawaCG's collision rows depend on (src/Solver.cpp:93,239,245; src/ConstraintSet.hpp:66)
		if( surface_inds.size() ){ ck( admm_b200_set_surface_inds( h, surface_inds.size(), surface_inds.data() ), "set_surface_inds" ); }
		ck( admm_b200_set_constraint_weight( h, m_constraints->constraint_w ), "set_constraint_weight" );
	}
	if( m_settings.linsolver == 1 ){
		ck( admm_b200_set_system( h, n, rowptr.data(), cols.data(), vals.data() ), "set_system" );
		// Replace the CPU solver by one whose colour lists can be read: GPU and cpu_step() then sweep the SAME colours
		std::shared_ptr<ColorPeek> gs = std::make_shared<ColorPeek>( m_constraints );
		gs->update_system( solver_termA );
		m_linsolver = gs;
		colors = gs->lists();
		std::vector<int> off(1,0), nodes;
		for( size_t c=0; c<colors.size(); ++c ){ nodes.insert( nodes.end(), colors[c].begin(), colors[c].end() ); off.push_back( nodes.size() ); }
		ck( admm_b200_set_colors( h, colors.size(), off.data(), nodes.data() ), "set_colors" );
		push_pins();
	} else {
		// LDLT / UzawaCG: the reference's own factorisation class on the n x n scalar matrix (LinearSolver.hpp:65-84);
		// its pieces are public: matrixL (unit lower, CSC, strictly lower entries stored), vectorD, permutationPinv
		for( int i=0; i<n; ++i ){
			if( !( m_masses[3*i]==m_masses[3*i+1] && m_masses[3*i]==m_masses[3*i+2] ) ){ throw std::runtime_error("**GpuSolver Error: LDLT needs equal x/y/z masses per node"); }
		}
		typedef Eigen::SparseMatrix<double> ColMat;
		std::vector< Eigen::Triplet<double> > at;
		for( int i=0; i<n; ++i ){ for( int q=rowptr[i]; q<rowptr[i+1]; ++q ){ at.emplace_back( i, cols[q], vals[q] + ( cols[q]==i ? m_masses[3*i] : 0.0 ) ); } }
		ColMat As(n,n); As.setFromTriplets( at.begin(), at.end() );
		Eigen::SimplicialLDLT<ColMat> chol( As );
		if( chol.info() != Eigen::Success ){ throw std::runtime_error("**GpuSolver Error: factorisation of the scalar system failed"); }
		const ColMat &L = chol.matrixL().nestedExpression();
		std::vector<int> Lp( L.outerIndexPtr(), L.outerIndexPtr()+n+1 ), perm(n);
		const int nnzL = Lp[n];
		std::vector<int> Li( L.innerIndexPtr(), L.innerIndexPtr()+nnzL );
		std::vector<double> Lx( L.valuePtr(), L.valuePtr()+nnzL ), D(n);
		for( int k=0; k<n; ++k ){ D[k] = chol.vectorD()[k]; perm[k] = chol.permutationPinv().indices()[k]; } // perm[new] = old
		ck( admm_b200_set_ldlt( h, n, perm.data(), Lp.data(), Li.data(), Lx.data(), D.data() ), "set_ldlt" );
	}

	// NodalMultiColorGS defaults (src/NodalMultiColorGS.hpp:45-46)
	ck( admm_b200_finalize( h, m_settings.timestep_s, m_settings.linsolver, 30, 1.9, 1e-10, precision ), "finalize" );
	if( admm_b200_pin_host( h, m_x.data(), sizeof(double)*dof ) == 0 ){
		if( admm_b200_pin_host( h, m_v.data(), sizeof(double)*dof ) == 0 ){ pinned = true; }
		else { admm_b200_unpin_host( h, m_x.data() ); }
	}
	return true;
}

inline void GpuSolver::step(){
	if( !initialized || !h ){ throw std::runtime_error("**GpuSolver::step Error: not initialized"); }
	m_runtime = RuntimeData();
	admm_b200_runtime rt;
	// Explicit forces (src/Solver.cpp:53-54): the reference's own ExplicitForce::project on the host arrays, which are about
	// to travel to the device anyway -- any subclass works and WindForce keeps the reference's exact (thread-order
	// dependent) arithmetic.  admm_b200_add_wind is the device form for callers that keep the state resident.
	const int n_ext_forces = ext_forces.size();
	for( int i=0; i<n_ext_forces; ++i ){ ext_forces[i]->project( m_settings.timestep_s, m_x, m_v, m_masses ); }
	ck( admm_b200_step_host( h, m_settings.admm_iters, m_settings.gravity, m_x.data(), m_v.data(), &rt ), "step" );
	m_runtime.global_ms = rt.global_ms; m_runtime.local_ms = rt.local_ms; m_runtime.collision_ms = rt.collision_ms; m_runtime.inner_iters = rt.inner_iters;
	if( m_settings.verbose > 0 ){ m_runtime.print(m_settings); }
}

// ConstraintSet::pins as the Gauss-Seidel sweep consults them (src/NodalMultiColorGS.hpp:111-117)
inline void GpuSolver::push_pins(){
	std::vector<int> idx; std::vector<double> pos;
	std::unordered_map<int,Vec3>::const_iterator it = m_constraints->pins.begin();
	for( ; it != m_constraints->pins.end(); ++it ){ idx.push_back( it->first ); pos.push_back( it->second[0] ); pos.push_back( it->second[1] ); pos.push_back( it->second[2] ); }
	ck( admm_b200_set_gs_pins( h, idx.size(), idx.data(), pos.data() ), "set_gs_pins" );
}

inline void GpuSolver::set_pins( const std::vector<int> &inds, const std::vector<Vec3> &points ){
	Solver::set_pins( inds, points );     // ConstraintSet::pins and the SpringPins' position / active state (src/Solver.cpp:113-157)
	if( !initialized || !h ){ return; }
	if( m_settings.linsolver == 1 ){ push_pins(); return; }
	// energy-based pins: the same vertices may move or be switched off (src/Solver.cpp:135-156)
	std::vector<double> pos; std::vector<unsigned char> act;
	for( size_t i=0; i<pin_order.size(); ++i ){
		std::unordered_map<int,Vec3>::const_iterator it = m_constraints->pins.find( pin_order[i] );
		const bool on = it != m_constraints->pins.end();
		act.push_back( on ? 1 : 0 );
		for( int j=0; j<3; ++j ){ pos.push_back( on ? it->second[j] : 0.0 ); }
	}
	ck( admm_b200_update_pins( h, pin_order.size(), pos.data(), act.data() ), "update_pins" );
}

} // end namespace admm

#endif
