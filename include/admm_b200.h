/* admm_b200.h -- C-ABI of the B200-native ADMM-elastic step.
 *
 * One opaque solver handle owns all device memory of ONE GPU.  Every entry point takes plain
 * pointers and sizes (host memory, borrowed for the duration of the call), returns 0 on success
 * and non-zero on failure with a message available from admm_b200_last_error().
 *
 * The library is a drop-in for the hot path of the reference (mattoverby/admm-elastic @ c6c09a3):
 * the body of admm::Solver::step() (src/Solver.cpp:35-110) -- per-element EnergyTerm::update
 * (src/EnergyTerm.hpp:130-140), the right-hand side assembly (src/Solver.cpp:98) and the global
 * solve LinearSolver::solve (src/LinearSolver.hpp:48, src/NodalMultiColorGS.hpp:60-146,
 * src/LinearSolver.hpp:87-90, src/UzawaCG.hpp:57-125).  The host keeps the reference's plugin
 * surface (admm::Solver / EnergyTerm / LinearSolver); what it harvests from that surface is
 * handed over through the calls below.  INTEGRATION.md shows the reference-side binding.
 *
 * There is NO CPU fallback: every compute entry point fails if no CUDA device is usable.
 */
#ifndef ADMM_B200_H
#define ADMM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct admm_b200_solver admm_b200_solver;

/* Constitutive model of a batch of tets.  Replaces the dynamic type of the reference's terms:
 * TetEnergyTerm (src/TetEnergyTerm.hpp:57), NeoHookeanTet (:117), StVKTet (:146), SplineTet (:179)
 * with the three built-in xu::Spline materials (src/XuSpline.hpp:48-94). */
enum admm_b200_tet_model {
	ADMM_B200_TET_LINEAR = 0,
	ADMM_B200_TET_NEOHOOKEAN = 1,
	ADMM_B200_TET_STVK = 2,
	ADMM_B200_TET_SPLINE_NH = 3,
	ADMM_B200_TET_SPLINE_STVK = 4,
	ADMM_B200_TET_SPLINE_COROT = 5
};

/* Settings::linsolver (src/Solver.hpp:46): 0=LDLT, 1=NodalMultiColorGS, 2=UzawaCG */
enum admm_b200_linsolver {
	ADMM_B200_LDLT = 0,
	ADMM_B200_MCGS = 1,
	ADMM_B200_UZAWA = 2
};

/* Storage/arithmetic type of per-element data (z, u, Dm^-1, corner forces).  Node data
 * (x, v, b) and the global solve are always fp64 on the device. */
enum admm_b200_precision {
	ADMM_B200_FP32 = 0,
	ADMM_B200_FP64 = 1
};

/* Passive obstacles handled inside the Gauss-Seidel sweep (src/PassiveObject.hpp:32-64). */
enum admm_b200_obstacle {
	ADMM_B200_FLOOR = 0,   /* params: y                      (Floor,  src/PassiveObject.hpp:32-45) */
	ADMM_B200_SPHERE = 1   /* params: cx, cy, cz, radius     (Sphere, src/PassiveObject.hpp:48-64) */
};

/* RuntimeData of the last step (src/Solver.hpp:54-61), times from CUDA events. */
typedef struct admm_b200_runtime {
	double global_ms;
	double local_ms;
	double collision_ms;
	int inner_iters;
	/* extra to the reference's fields: the b = M x_bar + D^T W (z-u) assembly share of global_ms,
	 * and the device time of the whole step (first to last kernel) */
	double assemble_ms;
	double step_ms;
} admm_b200_runtime;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Creates a solver on CUDA device `device`.  Fails (non-zero) when no CUDA device is present. */
int admm_b200_create( int device, admm_b200_solver **out );
void admm_b200_destroy( admm_b200_solver *s );
/* Message of the last failure on this handle (or of the last failed create when s == NULL). */
const char *admm_b200_last_error( const admm_b200_solver *s );
/* All work is enqueued on `cuda_stream` (a cudaStream_t; NULL = the solver's own stream). */
int admm_b200_set_stream( admm_b200_solver *s, void *cuda_stream );
int admm_b200_synchronize( admm_b200_solver *s );

/* ---- scene hand-over (before finalize) ----------------------------------------------------- */

/* Node state, replaces Solver::add_nodes / public m_x, m_v, m_masses (src/Solver.hpp:66-68,
 * 127-141).  x, m: 3*n_nodes doubles (xyz interleaved; masses "scaled x3" as in the reference).
 * v may be NULL (zeros; initialize() zeroes it, src/Solver.cpp:187). */
int admm_b200_set_nodes( admm_b200_solver *s, int n_nodes, const double *x, const double *v, const double *m );

/* A batch of n tets sharing one material.  Replaces what TetEnergyTerm's constructor and
 * get_reduction compute (src/TetEnergyTerm.cpp:31-71):
 *   idx[4e+c]          vertex c of tet e
 *   dminv[9e+3c+r]     edges_inv(c,r), the inverse rest edge matrix, so F = Ds * edges_inv
 *   weight[e]          get_weight() = sqrt(K*vol)
 *   row_offset[e]      g_index, the first of the tet's 9 rows of D (src/EnergyTerm.hpp:117);
 *                      only used to lay out debug_get("z"/"u") like the reference; may be NULL
 *   mu, lambda         constants of the constitutive model (for the spline models: the xu::Spline's own);
 *                      kappa: compression term of the spline models
 *   bulk_modulus       K of the prox penalty K/2 |sigma - sigma0|^2 = Lame::bulk_modulus() of the ELEMENT's Lame
 *                      (problem.k, src/TetEnergyTerm.hpp:125-128, 154-157, 193-200) -- it differs from the model constants
 *                      only for a SplineTet built with a spline of other constants; <= 0: lambda + 2/3 mu. */
int admm_b200_add_tets( admm_b200_solver *s, int n, const int *idx, const double *dminv, const double *weight,
	int model, double mu, double lambda, double kappa, double bulk_modulus, const int *row_offset );

/* A batch of n triangles (TriEnergyTerm, src/TriEnergyTerm.cpp:29-101):
 *   idx[3e+c]; restpose[4e+2c+r] = rest_pose(c,r) (2x2, F = [x1-x0, x2-x0] * rest_pose);
 *   weight[e] = sqrt(K*area); limit_min/limit_max = Lame::limit_min/max (src/EnergyTerm.hpp:46). */
int admm_b200_add_tris( admm_b200_solver *s, int n, const int *idx, const double *restpose, const double *weight,
	double limit_min, double limit_max, const int *row_offset );

/* Energy-based hard pins, SpringPin (src/SpringEnergyTerm.hpp:31-73), used with LDLT / Uzawa
 * (src/Solver.cpp:190-196): idx[i], pos[3i..], weight[i] = sqrt(2*K_rubber).  Each pin reserves
 * 6 rows of D of which 3 are live. */
int admm_b200_add_pins( admm_b200_solver *s, int n, const int *idx, const double *pos, const double *weight, const int *row_offset );
/* Per-frame update of the same pins: SpringPin::set_pin / set_active via Solver::set_pins
 * (src/Solver.cpp:135-156).  pin i of this call addresses the i-th pin added. */
int admm_b200_update_pins( admm_b200_solver *s, int n, const double *pos, const unsigned char *active );

/* Hard pins of the Gauss-Seidel solver: ConstraintSet::pins consulted inside the sweep
 * (src/NodalMultiColorGS.hpp:111-117).  May be called every frame with a different set. */
int admm_b200_set_gs_pins( admm_b200_solver *s, int n, const int *idx, const double *pos );

/* Passive obstacles tested per node inside the sweep, Collider::detect_passive
 * (src/Collider.hpp:137-150); order of calls = order of passive_objs. */
int admm_b200_add_obstacle( admm_b200_solver *s, int kind, const double *params );

/* Solver::ext_forces.push_back( WindForce(tris) ) (src/Solver.hpp:71, src/ExplicitForce.hpp:40-48): a wind force over a list
 * of triangles (3 vertex indices each), applied to the velocities at the top of every step, before gravity
 * (src/Solver.cpp:53-57; WindForce::project, src/ExplicitForce.cpp:47-104), in the order of the calls.  `direction` is
 * WindForce::direction and may be changed between steps.  Every kick is formed from the velocities before the call
 * (the reference's result depends on its thread count, see csrc/kernels.cuh: wind_tri_kernel).  Single-GPU. */
int admm_b200_add_wind( admm_b200_solver *s, const int *tris, int n_tris, const double *direction, int *id );
int admm_b200_set_wind_direction( admm_b200_solver *s, int id, const double *direction );

/* UzawaCG only.  The vertices Collider::detect tests for passive hits, in that order: Solver::surface_inds
 * (src/Solver.hpp:69, src/Solver.cpp:93, src/Collider.hpp:152-212; filled by binding::add_tetmesh,
 * samples/utils/AddMeshes.hpp:130-136).  n = 0 (the default): every node, in node order.  The order fixes the row order of
 * the constraint matrix and with it which multiplier a warm start hands to which row (src/UzawaCG.hpp:69-74). */
int admm_b200_set_surface_inds( admm_b200_solver *s, int n, const int *idx );
/* UzawaCG only.  ConstraintSet::constraint_w: C and c are scaled by sqrt(max(0, w)) (src/ConstraintSet.hpp:66,84-88);
 * 1 unless Settings::constraint_w > 0 (-ck) overrides it (src/Solver.cpp:239,245). */
int admm_b200_set_constraint_weight( admm_b200_solver *s, double constraint_w );

/* The constant global matrix A = M + dt^2 D^T W^2 D (src/Solver.cpp:226).  Every get_reduction of
 * the reference couples x-x, y-y, z-z only, so A = L (x) I3 plus a diagonal: the host passes the
 * n x n scalar matrix L in CSR (both triangles, WITHOUT the mass term) and the library adds the
 * per-component masses to the diagonal. */
int admm_b200_set_system( admm_b200_solver *s, int n, const int *rowptr, const int *cols, const double *vals );

/* Colour -> node lists computed once on the host, graphcolor::color_matrix(A, colors, 3)
 * (deps/mclscene/include/MCL/GraphColor.hpp:66-72) or any valid colouring. */
int admm_b200_set_colors( admm_b200_solver *s, int n_colors, const int *offsets, const int *nodes );

/* Prefactored L D L^T of P (L_scalar + M) P^T, replaces Cholesky::compute in
 * LDLTSolver::update_system (src/LinearSolver.hpp:79-84): perm[new] = old, unit lower L in CSC
 * without the diagonal (Lp, Li, Lx), D = diagonal.  Requires equal x/y/z masses per node. */
int admm_b200_set_ldlt( admm_b200_solver *s, int n, const int *perm, const int *Lp, const int *Li, const double *Lx, const double *D );

/* Ends the hand-over: Solver::initialize (src/Solver.cpp:167-261).  dt = Settings::timestep_s;
 * gs_iters, gs_omega, gs_tol = NodalMultiColorGS::max_iters, m_omega, m_tol
 * (src/NodalMultiColorGS.hpp:41-46; gs_tol <= 0 disables the residual test). */
int admm_b200_finalize( admm_b200_solver *s, double dt, int linsolver, int gs_iters, double gs_omega, double gs_tol, int precision );

/* ---- the hot path -------------------------------------------------------------------------- */

/* One Solver::step() (src/Solver.cpp:35-110) on device-resident state. runtime may be NULL. */
int admm_b200_step( admm_b200_solver *s, int admm_iters, double gravity, admm_b200_runtime *runtime );
/* The same with host state: uploads x, v (3n doubles each), steps, downloads them again -- what a
 * caller that reads m_x after every step() sees (samples/utils/Application.hpp:274-299). */
int admm_b200_step_host( admm_b200_solver *s, int admm_iters, double gravity, double *x, double *v, admm_b200_runtime *runtime );
int admm_b200_upload_state( admm_b200_solver *s, const double *x, const double *v );
/* Page-locks a caller-owned host range (the storage behind m_x / m_v, src/Solver.hpp:66-67) so the
 * per-step copies of admm_b200_step_host run at full PCIe/C2C rate and asynchronously; unpin before
 * the memory is freed.  Pinning is an optimisation only: step_host accepts pageable memory too. */
int admm_b200_pin_host( admm_b200_solver *s, void *ptr, unsigned long long bytes );
int admm_b200_unpin_host( admm_b200_solver *s, void *ptr );
int admm_b200_download_state( admm_b200_solver *s, double *x, double *v );

/* ---- pieces of the path, for parity tests and micro-benchmarks ----------------------------- */

/* prox() alone on n deformation gradients, column-major 9 doubles each, in the chosen precision:
 * TetEnergyTerm::prox / HyperElasticTet::prox (src/TetEnergyTerm.cpp:73-92, 114-136). */
int admm_b200_prox_tets( admm_b200_solver *s, int model, double mu, double lambda, double kappa, double bulk_modulus,
	int precision, int n, const double *z_in, double *z_out );
/* TriEnergyTerm::prox (src/TriEnergyTerm.cpp:73-101), 6 doubles each. */
int admm_b200_prox_tris( admm_b200_solver *s, double limit_min, double limit_max,
	int precision, int n, const double *z_in, double *z_out );
/* LinearSolver::solve alone: x (3n, in: warm start, out: solution), b (3n). Returns the solver's
 * iteration count in *iters (NodalMultiColorGS returns the sweep count, LDLT 1). */
int admm_b200_linsolve( admm_b200_solver *s, double *x, const double *b, int *iters );
/* Copies an internal array to the host: "z", "u" (reference row layout, needs row_offset),
 * "b", "x" (current iterate), "v".  n_out = capacity in doubles. */
int admm_b200_debug_get( admm_b200_solver *s, const char *name, double *out, long long n_out );
/* Keep z on the device (debug_get "z"); off by default because the path never re-reads z. */
int admm_b200_set_debug( admm_b200_solver *s, int store_z );

/* Device timing of the last `admm_b200_time_kernels` call: runs each kernel of one ADMM iteration
 * `reps` times back to back and reports the average milliseconds of local (tet prox), assemble and
 * global (solve) in out_ms[3].  Used by bench.py for the roofline numbers. */
int admm_b200_time_kernels( admm_b200_solver *s, int reps, double *out_ms );

/* Deferred timers: with `on`, admm_b200_step / _step_host called WITHOUT a runtime pointer still record their CUDA
 * events but do not synchronise, so the host can queue the next step while this one runs (a runtime pointer makes a
 * step wait for its own events).  admm_b200_collect_timers waits for the stream and returns the SUMS over all steps
 * since the last collection (RuntimeData fields summed; *steps = how many); admm_b200_kernel_times then holds the
 * kernel-only sums of the same steps.  on = n > 1 SAMPLES: only every n-th step (the first, the n+1-th, ...) records
 * events and is counted -- an event costs about 3 us of stream time and a step records 10 per ADMM iteration. */
int admm_b200_set_deferred_timers( admm_b200_solver *s, int on );
int admm_b200_collect_timers( admm_b200_solver *s, admm_b200_runtime *sum, int *steps );

/* Kernel-only device times of the last timed step (admm_b200_step* with a runtime pointer): CUDA events recorded on
 * the solver's stream immediately before and after each launch of the three hot kernels -- [0] tet prox kernel
 * (without its queue consumer), [1] assemble kernel, [2] solve kernel (without its scratch memset).  out_ms = summed
 * milliseconds, out_n = launches.  bench.py's roofline uses these; the RuntimeData phases include the helpers. */
int admm_b200_kernel_times( admm_b200_solver *s, double *out_ms /* [3] */, long long *out_n /* [3] */ );

/* Counts of kernels launched by this handle since creation (bench.py's gpu_launches). */
long long admm_b200_launch_count( const admm_b200_solver *s );

/* ---- multi-GPU: one handle per rank (process), at most 8 ranks on one NVLink box ----------------
 * Every rank is handed the SAME nodes, system matrix, colours and pins, but only the elements that touch
 * a node it owns (node -> part from admm_b200_plan_parts with n_parts = world * admm_b200_gs_parts();
 * owner rank = part / gs_parts).  Elements on a cut are therefore computed by both neighbours (a few per
 * cent), which makes the right-hand side assembly local.  The one exchange of the path -- neighbour
 * values inside the Gauss-Seidel sweeps and the solved positions of cut nodes -- is done by the solve
 * kernel itself with stores into the peers' memory (CUDA IPC mappings of their buffers); the 256-byte
 * blobs are moved between the processes by the caller (e.g. torch.distributed.all_gather).
 * Restrictions: NodalMultiColorGS, precision FP32; the reference's convergence test (which never fires)
 * is off, the sweep count is returned.  admm_b200_step_host / _upload_state / _download_state move only this rank's
 * nodes (see admm_b200_mgpu_nodes). */
#define ADMM_B200_IPC_BYTES 256
int admm_b200_set_rank( admm_b200_solver *s, int rank, int world );                /* before finalize */
int admm_b200_device_sms( const admm_b200_solver *s );
/* Parts (= CTAs, one per SM by default) of the shared-memory-resident Gauss-Seidel on this GPU.  Fewer, larger parts
 * are a testing aid: a 100k-tet mesh cut into 16 parts gives each part as many nodes, slices per warp, halo nodes and
 * neighbours as the 1M-tet bench mesh has with 148, at a size the CPU reference finishes in a second.  Before finalize;
 * 0 = one part per SM.  With several ranks the owner of a node is part / admm_b200_gs_parts(). */
int admm_b200_set_gs_parts( admm_b200_solver *s, int n_parts );
int admm_b200_gs_parts( const admm_b200_solver *s );
int admm_b200_plan_parts( int n, const int *rowptr, const int *cols, const double *vals, const double *pos3, int n_parts, int *part_of );
/* After finalize: how many nodes this rank owns and how many it reads from other ranks (ghosts).  With several ranks
 * admm_b200_step_host moves only those, in both directions (a ghost's value is the copy its owner pushed); the entries of
 * the caller's x / v arrays of all other nodes are left untouched (merge the ranks' arrays by owner). */
int admm_b200_mgpu_nodes( const admm_b200_solver *s, int *n_owned, int *n_ghost );
int admm_b200_mgpu_export( admm_b200_solver *s, void *blob );                      /* after finalize */
int admm_b200_mgpu_import( admm_b200_solver *s, int peer_rank, const void *blob );
int admm_b200_mgpu_ready( admm_b200_solver *s );
/* Host-only (tests): what rank `rank` would exchange -- per node the ranks that read it (owned nodes), a
 * ghost flag (nodes read from another rank) and the owning rank. */
int admm_b200_mgpu_plan_check( int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off, const int *color_nodes,
	const double *pos3, int sms, int world, int rank, unsigned int *mask_out, int *ghost_out, int *owner_out );

/* Host-only model of a barrier-free (slice-level dataflow) schedule of the resident Gauss-Seidel, the next step of
 * DESIGN.md 9.1 (csrc/dataflow_plan.hpp): random legal schedules must reproduce colour-by-colour SOR sweeps bit for
 * bit and never deadlock.  No kernel uses the schedule yet.  stats[4] = {max own-slice dependencies, max halo
 * references of a slice, tasks executed, parts used}. */
int admm_b200_dataflow_check( int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off, const int *color_nodes,
	const double *pos3, int n_parts, int n_warps, int sweeps, int n_seeds, long long *stats );

/* Host-only self check of the shared-memory-resident Gauss-Seidel plan (no device needed): see
 * csrc/partition.hpp.  Returns 0 when the plan covers every node exactly once and reproduces
 * L_offdiag * x; the message of a failure is available from admm_b200_last_error(NULL). */
int admm_b200_plan_check( int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off, const int *color_nodes,
	const double *pos3, int n_parts, int val_bytes, int lanes, const double *x, double *max_err, long long *stats, int *part_of );

/* Host-only: modelled shared-memory cycles of the resident sweep's float4 gathers (one sweep, all parts) with the row
 * entries in matrix order, after the bank-conflict-aware ordering of csrc/partition.hpp (detail::schedule_slice), and
 * the conflict-free minimum.  out[8] = {before, after, minimum, shared bytes / max slices / ELL fill x 1000 of the tiled
 * kernel's plan (4 lanes per node), shared bytes of the one-lane plan, max rows of the tiled plan}. */
int admm_b200_plan_bank_stats( int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off, const int *color_nodes,
	const double *pos3, int n_parts, long long *out );

/* Host-only: the block (supernodal) plan of the L D L^T solve (csrc/ldlt_blocks.hpp) for the factor of admm_b200_set_ldlt's
 * form, applied on the host to one right-hand side b (n values) exactly as the device kernel walks it.
 * stats[8] = {blocks, largest block, forward levels, backward levels, entries outside / inside the inverted diagonal blocks,
 * cut level of the bottom forest (levels below it need no grid barrier), forest segments}. */
int admm_b200_ldlt_blocks_check( int n, const int *perm, const int *Lp, const int *Li, const double *Lx, const double *D, const double *b, double *x, long long *stats );

/* One line describing which global-solve kernel finalize chose and why (diagnostics). */
const char *admm_b200_solver_info( const admm_b200_solver *s );

int admm_b200_version( void );

#ifdef __cplusplus
}
#endif
#endif
