// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" driver around the UNMODIFIED reference library
// (/root/reference/src/*.cpp, compiled where they lie by oracle/Makefile into
// oracle/_ref/libadmm_ref.so).  It exists so that tests, golden-vector
// generation and bench.py's CPU baseline can run the reference's own
// admm::Solver / EnergyTerm / LinearSolver code from Python (ctypes).
//
// Nothing here re-implements reference arithmetic: every number comes out of
// admm::Solver::initialize/step (src/Solver.cpp:35-261), EnergyTerm::update
// (src/EnergyTerm.hpp:130-140), the prox functions (src/TetEnergyTerm.cpp,
// src/TriEnergyTerm.cpp, src/SpringEnergyTerm.hpp) and the LinearSolvers
// (src/LinearSolver.hpp, src/NodalMultiColorGS.hpp, src/UzawaCG.hpp).
// Protected members are reached by subclassing, as SURVEY.md App. C describes.
//
// Only tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline
// legs may load this library.

#include "Solver.hpp"
#include "TetEnergyTerm.hpp"
#include "TriEnergyTerm.hpp"
#include "NodalMultiColorGS.hpp"
#include "UzawaCG.hpp"
#include "PassiveObject.hpp"
#include "MCL/TetMesh.hpp"
#include "MCL/MeshIO.hpp"
#include <cstring>
#include <chrono>

namespace {

using admm::Lame;
typedef Eigen::Matrix<double,Eigen::Dynamic,1> VecX;
typedef Eigen::SparseMatrix<double,Eigen::RowMajor> SparseMat;

// NodalMultiColorGS with its colour lists exposed (they are protected).
struct PeekGS : public admm::NodalMultiColorGS {
	PeekGS( std::shared_ptr<admm::ConstraintSet> c ) : admm::NodalMultiColorGS(c) {}
	std::vector< std::vector<int> > &colors(){ return A_colors; }
};

struct PeekSolver : public admm::Solver {
	std::shared_ptr<PeekGS> gs; // non-null when linsolver==1
	std::string error;

	// Same as Solver::initialize, then swap in a PeekGS that shares the
	// constraint set so that colours can be read / overridden.
	bool init( const Settings &s ){
		bool ok = admm::Solver::initialize(s);
		if( ok && m_settings.linsolver==1 ){
			gs = std::make_shared<PeekGS>( m_constraints );
			gs->update_system( solver_termA );
			m_linsolver = gs;
		}
		return ok;
	}
	void set_admm_iters( int it ){ m_settings.admm_iters = it; } // Settings are copied at initialize (src/Solver.cpp:168)
	const SparseMat &D() const { return m_D; }
	const VecX &W() const { return m_W_diag; }
	const SparseMat &A() const { return solver_termA; }
	const SparseMat &DtWtW() const { return solver_Dt_Wt_W; }
	admm::LinearSolver *linsolver(){ return m_linsolver.get(); }
	admm::ConstraintSet *constraints(){ return m_constraints.get(); }

	// One time step with every intermediate recorded.  This repeats the
	// statement order of Solver::step (src/Solver.cpp:35-110) using the
	// reference's own update()/solve() calls, so that z, u, b and x can be
	// captured after every ADMM iteration.  ext_forces are not applied.
	void traced_step( double *z_out, double *u_out, double *b_out, double *x_out ){
		const int dof = m_x.rows();
		const int n_nodes = dof/3;
		const double dt = m_settings.timestep_s;
		const int n_terms = energyterms.size();
		if( std::abs(m_settings.gravity)>0 ){
			for( int i=0; i<n_nodes; ++i ){ m_v[i*3+1] += dt*m_settings.gravity; }
		}
		VecX x_bar = m_x + dt * m_v;
		VecX M_xbar = m_masses.asDiagonal() * x_bar;
		VecX curr_x = x_bar;
		VecX curr_z = m_D*m_x;
		VecX curr_u = VecX::Zero( curr_z.rows() );
		VecX termB = VecX::Zero( dof );
		const int R = curr_z.rows();
		bool detect_passive = m_settings.linsolver!=1;
		for( int s_i=0; s_i < m_settings.admm_iters; ++s_i ){
			#pragma omp parallel for
			for( int i=0; i<n_terms; ++i ){ energyterms[i]->update( m_D, curr_x, curr_z, curr_u ); }
			m_constraints->collider->clear_hits();
			m_constraints->collider->detect( surface_inds, curr_x, detect_passive );
			termB.noalias() = M_xbar + solver_Dt_Wt_W * ( curr_z - curr_u );
			m_linsolver->solve( curr_x, termB );
			if( z_out ){ std::memcpy( z_out + (size_t)s_i*R, curr_z.data(), sizeof(double)*R ); }
			if( u_out ){ std::memcpy( u_out + (size_t)s_i*R, curr_u.data(), sizeof(double)*R ); }
			if( b_out ){ std::memcpy( b_out + (size_t)s_i*dof, termB.data(), sizeof(double)*dof ); }
			if( x_out ){ std::memcpy( x_out + (size_t)s_i*dof, curr_x.data(), sizeof(double)*dof ); }
		}
		m_v.noalias() = ( curr_x - m_x ) * ( 1.0 / dt );
		m_x = curr_x;
	}
};

struct Handle {
	PeekSolver solver;
	admm::Solver::Settings settings;
	std::vector<int> pin_inds;
	std::vector<Eigen::Vector3d> pin_pts;
};

Lame make_lame( double mu, double lambda, double lmin, double lmax ){
	Lame l; l.mu = mu; l.lambda = lambda; l.limit_min = lmin; l.limit_max = lmax; return l;
}

std::shared_ptr<admm::xu::Spline> make_spline( int kind, double mu, double lambda, double kappa ){
	switch( kind ){
		case 0: return std::make_shared<admm::xu::NeoHookean>( mu, lambda, kappa );
		case 1: return std::make_shared<admm::xu::StVK>( mu, lambda, kappa );
		default: return std::make_shared<admm::xu::CoRotated>( mu, lambda, kappa );
	}
}

template <typename F> int guarded( Handle *h, F f ){
	try { f(); return 0; }
	catch( std::exception &e ){ if(h){ h->solver.error = e.what(); } return 1; }
}

} // anon ns

extern "C" {

// model ids shared with include/admm_b200.h
enum { REF_TET_LINEAR=0, REF_TET_NEOHOOKEAN=1, REF_TET_STVK=2,
       REF_TET_SPLINE_NH=3, REF_TET_SPLINE_STVK=4, REF_TET_SPLINE_COROT=5 };

void *ref_create(){ return new Handle(); }
void ref_destroy( void *h ){ delete (Handle*)h; }
const char *ref_last_error( void *h ){ return ((Handle*)h)->solver.error.c_str(); }

int ref_add_nodes( void *h_, const double *x, const double *m, int n_verts ){
	Handle *h = (Handle*)h_;
	return h->solver.add_nodes( const_cast<double*>(x), const_cast<double*>(m), n_verts );
}

// verts/inds are mesh-local (inds index verts); vertex_offset shifts them into
// solver numbering exactly as create_tets_from_mesh does (src/TetEnergyTerm.hpp:35-51).
int ref_add_tets( void *h_, const double *verts, const int *inds, int n_tets, int model,
	double mu, double lambda, double kappa, int vertex_offset ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		Lame lame = make_lame( mu, lambda, -100.0, 100.0 );
		auto &et = h->solver.energyterms;
		switch( model ){
		case REF_TET_LINEAR: admm::create_tets_from_mesh<double,admm::TetEnergyTerm>( et, verts, inds, n_tets, lame, vertex_offset ); break;
		case REF_TET_NEOHOOKEAN: admm::create_tets_from_mesh<double,admm::NeoHookeanTet>( et, verts, inds, n_tets, lame, vertex_offset ); break;
		case REF_TET_STVK: admm::create_tets_from_mesh<double,admm::StVKTet>( et, verts, inds, n_tets, lame, vertex_offset ); break;
		default: {
			typedef Eigen::Matrix<int,4,1> Vec4i;
			typedef Eigen::Vector3d Vec3;
			for( int i=0; i<n_tets; ++i ){
				Vec4i tet( inds[i*4], inds[i*4+1], inds[i*4+2], inds[i*4+3] );
				std::vector<Vec3> tv;
				for( int c=0; c<4; ++c ){ tv.emplace_back( verts[tet[c]*3], verts[tet[c]*3+1], verts[tet[c]*3+2] ); }
				tet += Vec4i(1,1,1,1)*vertex_offset;
				et.emplace_back( std::make_shared<admm::SplineTet>( tet, tv, lame,
					make_spline( model-REF_TET_SPLINE_NH, mu, lambda, kappa ) ) );
			}
		} break;
		}
	});
}

// SplineTet whose spline has constants of its own (src/TetEnergyTerm.hpp:200-205); spline_type 0 NeoHookean, 1 StVK, 2 CoRotated
int ref_add_spline_tets( void *h_, const double *verts, const int *inds, int n_tets, int spline_type,
	double mu, double lambda, double sp_mu, double sp_lambda, double sp_kappa, int vertex_offset ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		typedef Eigen::Matrix<int,4,1> Vec4i; typedef Eigen::Vector3d Vec3;
		Lame lame = make_lame( mu, lambda, -100.0, 100.0 );
		for( int i=0; i<n_tets; ++i ){
			Vec4i tet( inds[i*4], inds[i*4+1], inds[i*4+2], inds[i*4+3] );
			std::vector<Vec3> tv;
			for( int c=0; c<4; ++c ){ tv.emplace_back( verts[tet[c]*3], verts[tet[c]*3+1], verts[tet[c]*3+2] ); }
			tet += Vec4i(1,1,1,1)*vertex_offset;
			h->solver.energyterms.emplace_back( std::make_shared<admm::SplineTet>( tet, tv, lame, make_spline( spline_type, sp_mu, sp_lambda, sp_kappa ) ) );
		}
	});
}

int ref_add_tris( void *h_, const double *verts, const int *inds, int n_tris,
	double mu, double lambda, double limit_min, double limit_max, int vertex_offset ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		Lame lame = make_lame( mu, lambda, limit_min, limit_max );
		admm::create_tris_from_mesh<double,admm::TriEnergyTerm>( h->solver.energyterms, verts, inds, n_tris, lame, vertex_offset );
	});
}

// points==NULL pins in place (src/Solver.cpp:113-130)
int ref_set_pins( void *h_, const int *inds, const double *points, int n ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		std::vector<int> i( inds, inds+n );
		std::vector<Eigen::Vector3d> p;
		if( points ){ for( int k=0; k<n; ++k ){ p.emplace_back( points[3*k], points[3*k+1], points[3*k+2] ); } }
		h->solver.set_pins( i, p );
	});
}

// Solver::surface_inds (src/Solver.hpp:69): the vertices Collider::detect tests (src/Solver.cpp:93)
void ref_set_surface_inds( void *h_, const int *inds, int n ){ ((Handle*)h_)->solver.surface_inds.assign( inds, inds+n ); }

// Solver::ext_forces (src/Solver.hpp:71) with the reference's own WindForce (src/ExplicitForce.hpp:40-48)
int ref_add_wind( void *h_, const int *tris, int n_tris, const double *dir ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		std::vector<int> t( tris, tris+3*n_tris );
		std::shared_ptr<admm::WindForce> w = std::make_shared<admm::WindForce>( t );
		w->direction = Eigen::Vector3d( dir[0], dir[1], dir[2] );
		h->solver.ext_forces.emplace_back( w );
	} );
}
// WindForce::project alone on caller-owned arrays (src/ExplicitForce.cpp:47-104)
void ref_wind_project( const int *tris, int n_tris, const double *dir, double dt, int n_nodes, const double *x, double *v ){
	std::vector<int> t( tris, tris+3*n_tris );
	admm::WindForce w( t );
	w.direction = Eigen::Vector3d( dir[0], dir[1], dir[2] );
	Eigen::VectorXd xx = Eigen::Map<const Eigen::VectorXd>( x, 3*n_nodes ), vv = Eigen::Map<const Eigen::VectorXd>( v, 3*n_nodes ), mm = Eigen::VectorXd::Ones( 3*n_nodes );
	w.project( dt, xx, vv, mm );
	Eigen::Map<Eigen::VectorXd>( v, 3*n_nodes ) = vv;
}

int ref_add_floor( void *h_, double y ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){ h->solver.add_obstacle( std::make_shared<admm::Floor>( y ) ); } );
}

int ref_add_sphere( void *h_, const double *c, double r ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){ h->solver.add_obstacle( std::make_shared<admm::Sphere>( Eigen::Vector3d(c[0],c[1],c[2]), r ) ); } );
}

// returns 0 ok, 1 exception, 2 initialize()==false
int ref_initialize( void *h_, double dt, int admm_iters, double gravity, int linsolver, double constraint_w ){
	Handle *h = (Handle*)h_;
	int rc = 0;
	int e = guarded( h, [&](){
		admm::Solver::Settings s;
		s.timestep_s = dt; s.verbose = 0; s.admm_iters = admm_iters;
		s.gravity = gravity; s.linsolver = linsolver; s.constraint_w = constraint_w;
		h->settings = s;
		if( !h->solver.init( s ) ){ rc = 2; }
	});
	return e ? e : rc;
}

void ref_set_admm_iters( void *h_, int it ){ ((Handle*)h_)->solver.set_admm_iters( it ); }

int ref_step( void *h_ ){ Handle *h=(Handle*)h_; return guarded( h, [&](){ h->solver.step(); } ); }

int ref_traced_step( void *h_, double *z, double *u, double *b, double *x ){
	Handle *h=(Handle*)h_; return guarded( h, [&](){ h->solver.traced_step( z, u, b, x ); } );
}

// runtime[4] = global_ms, local_ms, collision_ms, inner_iters (src/Solver.hpp:54-61)
void ref_runtime( void *h_, double *out ){
	const admm::Solver::RuntimeData &r = ((Handle*)h_)->solver.runtime_data();
	out[0]=r.global_ms; out[1]=r.local_ms; out[2]=r.collision_ms; out[3]=r.inner_iters;
}

int ref_dof( void *h_ ){ return ((Handle*)h_)->solver.m_x.rows(); }
int ref_n_terms( void *h_ ){ return ((Handle*)h_)->solver.energyterms.size(); }
void ref_get_x( void *h_, double *x ){ auto &s=((Handle*)h_)->solver; std::memcpy( x, s.m_x.data(), sizeof(double)*s.m_x.rows() ); }
void ref_get_v( void *h_, double *v ){ auto &s=((Handle*)h_)->solver; std::memcpy( v, s.m_v.data(), sizeof(double)*s.m_v.rows() ); }
void ref_set_x( void *h_, const double *x ){ auto &s=((Handle*)h_)->solver; std::memcpy( s.m_x.data(), x, sizeof(double)*s.m_x.rows() ); }
void ref_set_v( void *h_, const double *v ){ auto &s=((Handle*)h_)->solver; std::memcpy( s.m_v.data(), v, sizeof(double)*s.m_v.rows() ); }

// Sparse dumps: which = 0:D  1:A  2:dt^2 D^T W^2.  CSR, row-major as the reference stores them.
static const SparseMat &pick( Handle *h, int which ){
	return which==0 ? h->solver.D() : ( which==1 ? h->solver.A() : h->solver.DtWtW() );
}
void ref_sparse_shape( void *h_, int which, long long *out ){
	const SparseMat &M = pick( (Handle*)h_, which );
	out[0]=M.rows(); out[1]=M.cols(); out[2]=M.nonZeros();
}
void ref_sparse_get( void *h_, int which, int *rowptr, int *cols, double *vals ){
	SparseMat M = pick( (Handle*)h_, which ); M.makeCompressed();
	std::memcpy( rowptr, M.outerIndexPtr(), sizeof(int)*(M.rows()+1) );
	std::memcpy( cols, M.innerIndexPtr(), sizeof(int)*M.nonZeros() );
	std::memcpy( vals, M.valuePtr(), sizeof(double)*M.nonZeros() );
}
int ref_n_weights( void *h_ ){ return ((Handle*)h_)->solver.W().rows(); }
void ref_get_weights( void *h_, double *w ){ const VecX &W=((Handle*)h_)->solver.W(); std::memcpy( w, W.data(), sizeof(double)*W.rows() ); }

// Colours of NodalMultiColorGS (linsolver 1 only). n_colors, then flattened lists.
int ref_n_colors( void *h_ ){ Handle *h=(Handle*)h_; return h->solver.gs ? (int)h->solver.gs->colors().size() : -1; }
void ref_get_colors( void *h_, int *offsets, int *nodes ){
	auto &c = ((Handle*)h_)->solver.gs->colors();
	int k=0; offsets[0]=0;
	for( size_t i=0; i<c.size(); ++i ){
		for( size_t j=0; j<c[i].size(); ++j ){ nodes[k++] = c[i][j]; }
		offsets[i+1]=k;
	}
}
void ref_set_colors( void *h_, int n_colors, const int *offsets, const int *nodes ){
	auto &c = ((Handle*)h_)->solver.gs->colors();
	c.clear();
	for( int i=0; i<n_colors; ++i ){ c.emplace_back( nodes+offsets[i], nodes+offsets[i+1] ); }
}
// GS knobs (src/NodalMultiColorGS.hpp:41-46)
void ref_gs_params( void *h_, int max_iters, double tol, double omega ){
	auto gs = ((Handle*)h_)->solver.gs; if(!gs){ return; }
	gs->max_iters = max_iters; gs->m_tol = tol; gs->m_omega = omega;
}

// Global solve alone: x (in/out, warm start) and b; returns the solver's return value.
int ref_linsolve( void *h_, double *x, const double *b ){
	Handle *h=(Handle*)h_; int dof = h->solver.m_x.rows(); int it=-1;
	guarded( h, [&](){
		VecX xv = Eigen::Map<VecX>( x, dof ); VecX bv = Eigen::Map<const VecX>( b, dof );
		it = h->solver.linsolver()->solve( xv, bv );
		std::memcpy( x, xv.data(), sizeof(double)*dof );
	});
	return it;
}

// Stand-alone prox evaluation on n deformation-gradient vectors (9 per tet,
// 6 per tri, column-major as EnergyTerm::update hands them to prox()).
// The element geometry is a unit right tet / tri: the prox result does not
// depend on it because w^2 = K*vol (src/TetEnergyTerm.cpp:47,88).
int ref_prox_tets( int model, double mu, double lambda, double kappa, int n, const double *z_in, double *z_out ){
	typedef Eigen::Matrix<int,4,1> Vec4i; typedef Eigen::Vector3d Vec3;
	try {
		Lame lame = make_lame( mu, lambda, -100.0, 100.0 );
		std::vector<Vec3> tv = { Vec3(0,0,0), Vec3(0,1,0), Vec3(0,0,1), Vec3(1,0,0) };
		Vec4i tet(0,1,2,3);
		std::shared_ptr<admm::TetEnergyTerm> t;
		switch( model ){
			case REF_TET_LINEAR: t = std::make_shared<admm::TetEnergyTerm>( tet, tv, lame ); break;
			case REF_TET_NEOHOOKEAN: t = std::make_shared<admm::NeoHookeanTet>( tet, tv, lame ); break;
			case REF_TET_STVK: t = std::make_shared<admm::StVKTet>( tet, tv, lame ); break;
			default: t = std::make_shared<admm::SplineTet>( tet, tv, lame, make_spline( model-REF_TET_SPLINE_NH, mu, lambda, kappa ) ); break;
		}
		for( int i=0; i<n; ++i ){
			VecX zi = Eigen::Map<const VecX>( z_in+9*i, 9 );
			t->prox( zi );
			std::memcpy( z_out+9*i, zi.data(), sizeof(double)*9 );
		}
	} catch( std::exception &e ){ fprintf( stderr, "ref_prox_tets: %s\n", e.what() ); return 1; }
	return 0;
}

// SplineTet::prox with the element's Lame (mu, lambda -> the penalty K) and a spline of other constants
int ref_prox_spline_tets( int spline_type, double mu, double lambda, double sp_mu, double sp_lambda, double sp_kappa, int n, const double *z_in, double *z_out ){
	typedef Eigen::Matrix<int,4,1> Vec4i; typedef Eigen::Vector3d Vec3;
	try {
		Lame lame = make_lame( mu, lambda, -100.0, 100.0 );
		std::vector<Vec3> tv = { Vec3(0,0,0), Vec3(0,1,0), Vec3(0,0,1), Vec3(1,0,0) };
		admm::SplineTet t( Vec4i(0,1,2,3), tv, lame, make_spline( spline_type, sp_mu, sp_lambda, sp_kappa ) );
		for( int i=0; i<n; ++i ){
			VecX zi = Eigen::Map<const VecX>( z_in+9*i, 9 );
			t.prox( zi );
			std::memcpy( z_out+9*i, zi.data(), sizeof(double)*9 );
		}
	} catch( std::exception &e ){ fprintf( stderr, "ref_prox_spline_tets: %s\n", e.what() ); return 1; }
	return 0;
}

int ref_prox_tris( double mu, double lambda, double limit_min, double limit_max, int n, const double *z_in, double *z_out ){
	typedef Eigen::Matrix<int,3,1> Vec3i; typedef Eigen::Vector3d Vec3;
	try {
		Lame lame = make_lame( mu, lambda, limit_min, limit_max );
		std::vector<Vec3> tv = { Vec3(0,0,0), Vec3(1,0,0), Vec3(0,1,0) };
		admm::TriEnergyTerm t( Vec3i(0,1,2), tv, lame );
		for( int i=0; i<n; ++i ){
			VecX zi = Eigen::Map<const VecX>( z_in+6*i, 6 );
			t.prox( zi );
			std::memcpy( z_out+6*i, zi.data(), sizeof(double)*6 );
		}
	} catch( std::exception &e ){ fprintf( stderr, "ref_prox_tris: %s\n", e.what() ); return 1; }
	return 0;
}

// Energy of one tet term on F (used by the test_lineartet known answers).
double ref_tet_energy( int model, double mu, double lambda, const double *verts12, const double *x12 ){
	typedef Eigen::Matrix<int,4,1> Vec4i; typedef Eigen::Vector3d Vec3;
	Lame lame = make_lame( mu, lambda, -100.0, 100.0 );
	std::vector<Vec3> tv; for( int c=0; c<4; ++c ){ tv.emplace_back( verts12[3*c], verts12[3*c+1], verts12[3*c+2] ); }
	std::shared_ptr<admm::EnergyTerm> t;
	Vec4i tet(0,1,2,3);
	if( model==REF_TET_LINEAR ){ t = std::make_shared<admm::TetEnergyTerm>( tet, tv, lame ); }
	else if( model==REF_TET_NEOHOOKEAN ){ t = std::make_shared<admm::NeoHookeanTet>( tet, tv, lame ); }
	else { t = std::make_shared<admm::StVKTet>( tet, tv, lame ); }
	std::vector< Eigen::Triplet<double> > trips; std::vector<double> w;
	t->get_reduction( trips, w );
	SparseMat D( 9, 12 ); D.setFromTriplets( trips.begin(), trips.end() );
	VecX x = Eigen::Map<const VecX>( x12, 12 );
	return t->energy( D, x );
}

// mcl::meshio::load_elenode (deps/mclscene/include/MCL/MeshIO.hpp:180-311) + TetMesh::weighted_masses / surface_inds:
// what the mesh loader of the package is checked against.  Returns a mesh handle (NULL on failure).
void *ref_mesh_load_elenode( const char *prefix ){
	mcl::TetMesh *m = new mcl::TetMesh();
	try { if( !mcl::meshio::load_elenode( m, std::string(prefix) ) ){ delete m; return NULL; } }
	catch( std::exception & ){ delete m; return NULL; }
	return m;
}
void ref_mesh_free( void *m ){ delete (mcl::TetMesh*)m; }
int ref_mesh_n_verts( void *m ){ return ((mcl::TetMesh*)m)->vertices.size(); }
int ref_mesh_n_tets( void *m ){ return ((mcl::TetMesh*)m)->tets.size(); }
void ref_mesh_get( void *m_, float *verts, int *tets, float *masses, float density ){
	mcl::TetMesh *m = (mcl::TetMesh*)m_;
	for( size_t i=0; i<m->vertices.size(); ++i ){ for( int j=0; j<3; ++j ){ verts[3*i+j] = m->vertices[i][j]; } }
	for( size_t i=0; i<m->tets.size(); ++i ){ for( int j=0; j<4; ++j ){ tets[4*i+j] = m->tets[i][j]; } }
	std::vector<float> w; m->weighted_masses( w, density );
	for( size_t i=0; i<w.size(); ++i ){ masses[i] = w[i]; }
}
int ref_mesh_surface_inds( void *m_, int *out ){
	std::vector<int> s; ((mcl::TetMesh*)m_)->surface_inds( s );
	if( out ){ for( size_t i=0; i<s.size(); ++i ){ out[i] = s[i]; } }
	return s.size();
}

int ref_omp_threads(){ return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1 to its workers: bench.py's reference arm asks for the host's cores again
void ref_set_omp_threads( int n ){ if( n > 0 ){ omp_set_num_threads(n); } }

} // extern C
