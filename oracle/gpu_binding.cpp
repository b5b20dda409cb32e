// oracle/gpu_binding.cpp -- TEST INFRASTRUCTURE: the reference-side binding (integration/GpuSolver.hpp, a subclass of
// the reference's admm::Solver that steps on the GPU through include/admm_b200.h) compiled against the UNMODIFIED
// reference headers where they lie under /root/reference, plus two ways to exercise it:
//
//  1. the reference's own test program, samples/tests/test_lineartet.cpp, compiled UNCHANGED from its place in the
//     reference tree with `admm::Solver` spelled `admm::GpuSolver` (a #define after every reference header has been
//     included once): its three tests -- energy / F-layout known answers, x = 52.2321 +- 1e-4 for 21..99 ADMM iterations
//     (:165-230), inversion recovery (:236-323) -- then run through initialize()/step() of the binding.
//     gpub_run_reference_lineartet_test() returns its exit code.
//  2. an extern "C" scene API (gpub_*) mirroring oracle/ref_harness.cpp so the Python parity tests can build a scene
//     with the reference's own create_tets_from_mesh / create_tris_from_mesh, step it on the GPU (step) or with the
//     reference's CPU code on the SAME object (cpu_step), and read what the binding harvested from get_reduction.
//
// Built by oracle/Makefile (target gpu_binding) into oracle/_ref/libadmm_gpubinding.so; links libadmm_b200.so.
#include "../integration/GpuSolver.hpp"
#include "MCL/Vec.hpp"
#include "MCL/XForm.hpp"
#include <cstring>
#include <iostream>

// ---- 1. the reference's test program, verbatim, against the binding -------------------------------------------------
#define Solver GpuSolver
#define main gpub_reference_lineartet_main
#include "samples/tests/test_lineartet.cpp"
#undef main
#undef Solver

// ---- 2. scene API for the Python tests ---------------------------------------------------------------------------------
namespace {
struct Handle {
	admm::GpuSolver solver;
	std::string error;
};
template <typename F> int guarded( Handle *h, F f ){
	try { f(); return 0; }
	catch( std::exception &e ){ if(h){ h->error = e.what(); } return 1; }
}
admm::Lame make_lame( double mu, double lambda, double lmin, double lmax ){
	admm::Lame l; l.mu = mu; l.lambda = lambda; l.limit_min = lmin; l.limit_max = lmax; return l;
}
}

extern "C" {

int gpub_run_reference_lineartet_test(){ return gpub_reference_lineartet_main(); }

void *gpub_create(){ return new Handle(); }
void gpub_destroy( void *h ){ delete (Handle*)h; }
const char *gpub_last_error( void *h ){ return ((Handle*)h)->error.c_str(); }
void gpub_set_options( void *h_, int precision, int gs_parts, int keep_z ){
	Handle *h = (Handle*)h_; h->solver.precision = precision; h->solver.gs_parts = gs_parts; h->solver.keep_z = keep_z != 0;
}

int gpub_add_nodes( void *h_, const double *x, const double *m, int n_verts ){
	return ((Handle*)h_)->solver.add_nodes( const_cast<double*>(x), const_cast<double*>(m), n_verts );
}

// model ids of include/admm_b200.h; spline models take their own (mu, lambda, kappa), the element its Lame
int gpub_add_tets( void *h_, const double *verts, const int *inds, int n_tets, int model, double mu, double lambda,
	double sp_mu, double sp_lambda, double sp_kappa, int vertex_offset ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		admm::Lame lame = make_lame( mu, lambda, -100.0, 100.0 );
		auto &et = h->solver.energyterms;
		switch( model ){
		case ADMM_B200_TET_LINEAR: admm::create_tets_from_mesh<double,admm::TetEnergyTerm>( et, verts, inds, n_tets, lame, vertex_offset ); break;
		case ADMM_B200_TET_NEOHOOKEAN: admm::create_tets_from_mesh<double,admm::NeoHookeanTet>( et, verts, inds, n_tets, lame, vertex_offset ); break;
		case ADMM_B200_TET_STVK: admm::create_tets_from_mesh<double,admm::StVKTet>( et, verts, inds, n_tets, lame, vertex_offset ); break;
		default: {
			typedef Eigen::Matrix<int,4,1> Vec4i; typedef Eigen::Vector3d Vec3;
			std::shared_ptr<admm::xu::Spline> sp;
			if( model == ADMM_B200_TET_SPLINE_NH ){ sp = std::make_shared<admm::xu::NeoHookean>( sp_mu, sp_lambda, sp_kappa ); }
			else if( model == ADMM_B200_TET_SPLINE_STVK ){ sp = std::make_shared<admm::xu::StVK>( sp_mu, sp_lambda, sp_kappa ); }
			else { sp = std::make_shared<admm::xu::CoRotated>( sp_mu, sp_lambda, sp_kappa ); }
			for( int i=0; i<n_tets; ++i ){
				Vec4i tet( inds[i*4], inds[i*4+1], inds[i*4+2], inds[i*4+3] );
				std::vector<Vec3> tv;
				for( int c=0; c<4; ++c ){ tv.emplace_back( verts[tet[c]*3], verts[tet[c]*3+1], verts[tet[c]*3+2] ); }
				tet += Vec4i(1,1,1,1)*vertex_offset;
				et.emplace_back( std::make_shared<admm::SplineTet>( tet, tv, lame, sp ) );
			}
		} break;
		}
	});
}

int gpub_add_tris( void *h_, const double *verts, const int *inds, int n_tris, double mu, double lambda, double limit_min, double limit_max, int vertex_offset ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		admm::Lame lame = make_lame( mu, lambda, limit_min, limit_max );
		admm::create_tris_from_mesh<double,admm::TriEnergyTerm>( h->solver.energyterms, verts, inds, n_tris, lame, vertex_offset );
	});
}

int gpub_set_pins( void *h_, const int *inds, const double *points, int n ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		std::vector<int> i( inds, inds+n );
		std::vector<Eigen::Vector3d> p;
		if( points ){ for( int k=0; k<n; ++k ){ p.emplace_back( points[3*k], points[3*k+1], points[3*k+2] ); } }
		h->solver.set_pins( i, p );
	});
}

int gpub_add_floor( void *h_, double y ){ Handle *h = (Handle*)h_; return guarded( h, [&](){ h->solver.add_obstacle( std::make_shared<admm::Floor>( y ) ); } ); }
int gpub_add_sphere( void *h_, const double *c, double r ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){ h->solver.add_obstacle( std::make_shared<admm::Sphere>( Eigen::Vector3d(c[0],c[1],c[2]), r ) ); } );
}

int gpub_add_wind( void *h_, const int *tris, int n_tris, const double *dir ){
	Handle *h = (Handle*)h_;
	return guarded( h, [&](){
		std::vector<int> t( tris, tris+3*n_tris );
		std::shared_ptr<admm::WindForce> w = std::make_shared<admm::WindForce>( t );
		w->direction = Eigen::Vector3d( dir[0], dir[1], dir[2] );
		h->solver.ext_forces.emplace_back( w );
	} );
}

// 0 ok, 1 exception, 2 initialize() == false
int gpub_initialize( void *h_, double dt, int admm_iters, double gravity, int linsolver ){
	Handle *h = (Handle*)h_;
	int rc = 0;
	int e = guarded( h, [&](){
		admm::Solver::Settings s;
		s.timestep_s = dt; s.verbose = 0; s.admm_iters = admm_iters; s.gravity = gravity; s.linsolver = linsolver;
		if( !h->solver.initialize( s ) ){ rc = 2; }
	});
	return e ? e : rc;
}

int gpub_step( void *h_ ){ Handle *h = (Handle*)h_; return guarded( h, [&](){ h->solver.step(); } ); }          // GPU
int gpub_cpu_step( void *h_ ){ Handle *h = (Handle*)h_; return guarded( h, [&](){ h->solver.cpu_step(); } ); }  // the reference's Solver::step()

int gpub_dof( void *h_ ){ return ((Handle*)h_)->solver.m_x.rows(); }
void gpub_get_x( void *h_, double *x ){ auto &s=((Handle*)h_)->solver; std::memcpy( x, s.m_x.data(), sizeof(double)*s.m_x.rows() ); }
void gpub_get_v( void *h_, double *v ){ auto &s=((Handle*)h_)->solver; std::memcpy( v, s.m_v.data(), sizeof(double)*s.m_v.rows() ); }
void gpub_set_x( void *h_, const double *x ){ auto &s=((Handle*)h_)->solver; std::memcpy( s.m_x.data(), x, sizeof(double)*s.m_x.rows() ); }
void gpub_set_v( void *h_, const double *v ){ auto &s=((Handle*)h_)->solver; std::memcpy( s.m_v.data(), v, sizeof(double)*s.m_v.rows() ); }
void gpub_runtime( void *h_, double *out ){
	const admm::Solver::RuntimeData &r = ((Handle*)h_)->solver.runtime_data();
	out[0]=r.global_ms; out[1]=r.local_ms; out[2]=r.collision_ms; out[3]=r.inner_iters;
}
const char *gpub_solver_info( void *h_ ){ return admm_b200_solver_info( ((Handle*)h_)->solver.handle() ); }

// What the binding harvested from the reference's plugin surface: all tet batches concatenated.
int gpub_n_tets( void *h_ ){ int n=0; for( auto &b : ((Handle*)h_)->solver.tet_batches ){ n += b.w.size(); } return n; }
void gpub_get_tets( void *h_, int *idx4, double *dminv9, double *w, int *row, int *model ){
	size_t e = 0;
	for( auto &b : ((Handle*)h_)->solver.tet_batches ){
		for( size_t i=0; i<b.w.size(); ++i, ++e ){
			for( int c=0; c<4; ++c ){ idx4[4*e+c] = b.idx[4*i+c]; }
			for( int k=0; k<9; ++k ){ dminv9[9*e+k] = b.dminv[9*i+k]; }
			w[e] = b.w[i]; row[e] = b.row[i]; model[e] = b.model;
		}
	}
}
int gpub_n_tris( void *h_ ){ int n=0; for( auto &b : ((Handle*)h_)->solver.tri_batches ){ n += b.w.size(); } return n; }
void gpub_get_tris( void *h_, int *idx3, double *rest4, double *w, int *row ){
	size_t e = 0;
	for( auto &b : ((Handle*)h_)->solver.tri_batches ){
		for( size_t i=0; i<b.w.size(); ++i, ++e ){
			for( int c=0; c<3; ++c ){ idx3[3*e+c] = b.idx[3*i+c]; }
			for( int k=0; k<4; ++k ){ rest4[4*e+k] = b.rest[4*i+k]; }
			w[e] = b.w[i]; row[e] = b.row[i];
		}
	}
}
int gpub_n_pins( void *h_ ){ return ((Handle*)h_)->solver.pin_idx.size(); }
void gpub_get_pins( void *h_, int *idx, int *row, double *w ){
	auto &s = ((Handle*)h_)->solver;
	for( size_t i=0; i<s.pin_idx.size(); ++i ){ idx[i] = s.pin_idx[i]; row[i] = s.pin_row[i]; w[i] = s.pin_w[i]; }
}
int gpub_n_colors( void *h_ ){ return ((Handle*)h_)->solver.colors.size(); }
void gpub_get_colors( void *h_, int *offsets, int *nodes ){
	auto &c = ((Handle*)h_)->solver.colors;
	int k=0; offsets[0]=0;
	for( size_t i=0; i<c.size(); ++i ){ for( size_t j=0; j<c[i].size(); ++j ){ nodes[k++] = c[i][j]; } offsets[i+1]=k; }
}
// z / u of the last iteration in the reference's row layout (needs keep_z), b, x
int gpub_debug_get( void *h_, const char *name, double *out, long long n ){
	Handle *h = (Handle*)h_;
	int rc = admm_b200_debug_get( h->solver.handle(), name, out, n );
	if( rc ){ h->error = admm_b200_last_error( h->solver.handle() ); }
	return rc;
}

} // extern C
