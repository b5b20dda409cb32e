// mcgs_resident_f32.cuh -- NodalMultiColorGS::solve (src/NodalMultiColorGS.hpp:60-146) on the INCREMENT,
// resident in shared memory, single precision sweeps around a double precision anchor.
//
// SOR on A x = b started at x_ref is, in exact arithmetic, the same iteration as SOR on
//     A d = r0,   r0 = b - A x_ref,   x = x_ref + d,   d = 0 at the start
// (substitute x = x_ref + d in segment_update, src/NodalMultiColorGS.hpp:180-215).  x_ref -- the previous
// ADMM iterate, metres -- stays in fp64 in HBM and is touched twice per solve (r0 at the start, x = x_ref
// + d at the end); the increment d -- millimetres within one solve -- is swept in fp32.  An fp32 ulp of d is
// 1e-10 m, i.e. better than 1e-10 relative to x: far below the 1e-6 the reference's own prox defines x
// to, while every neighbour read is one 16-byte float4 from shared memory instead of three
// doubles and all arithmetic is FFMA.  This is the production (ADMM_B200_FP32) global solve; the fp64
// variant (mcgs_resident_kernel) stays the validation path.
//
// Layout per part (one CTA per SM, partition.hpp mode 1): float4 d[own + halo], float val / u16 col sliced
// ELL, id tables.  Halo increments are cached in shared memory and refreshed once per pass -- only the
// halo nodes of the colour the neighbours have just updated (contiguous: halo is sorted by colour) --
// so the gather itself never leaves the SM.  Synchronisation, warp roles and the lazy convergence test
// are those of mcgs_resident.cuh.
#pragma once
#include "mcgs_resident.cuh"

namespace admmb200 {

#define ADMMB200_MAX_RANKS 8

struct McgsRes32Params {
	McgsParams base;
	const PartDesc *parts;
	const uint16_t *col;
	const float *val;     // swept in shared memory
	const double *val64;  // same entries in full precision, streamed once per solve for r0 = b - A x_ref
	const int *gid, *slice_row, *color_slice, *nbr, *halo_color;
	const short *slice_node;
	unsigned int *part_epoch, *sweep_flag, *sweep_arrive;
	unsigned long long *prof;
	uint2 *dglob;    // [2][n_nodes][3]: published increments of boundary nodes, each component one 64-bit word
	                 // {float bits, tag}; double-buffered by sweep parity
	float4 *nodebuf; // [2 * n_nodes] by (own_off + local id): {r0.xyz, pin slot + 1}, {1/a.xyz, -}
	unsigned int tag_base; // (solve sequence number << 12): tags never repeat across solves
	int n_nodes_total;
	// multi-GPU (world > 1): this rank runs parts [part0, part0 + gridDim.x) of a plan that spans all ranks.
	// Boundary values are also pushed into the peers that read them (stores over NVLink into their dglob /
	// x arrays, mapped through CUDA IPC); nothing is ever READ from a peer, so spinning stays local.
	int part0, world, rank;
	// mailboxes of the static-ownership kernel (partition.hpp, plan_mailboxes); dglob is then [2][3][total_slots]
	const int *dest_off; const unsigned int *dest_slot; int total_slots;
	int dbg; // timing experiments only (ADMM_B200_GS_DBG, tools/gs_prof.py): 1 no tag wait, 2 no refresh, 4 no interior, 8 no boundary (owned kernel: 2, 4, 16 no publish, 32 no gather, 64 fence after publish; bits 8.. = part whose passes 40..43 are traced)
	const unsigned int *dest_mask; // [n_nodes] ranks (bit q) that read this node as halo; NULL when world == 1
	uint2 *peer_dglob[ADMMB200_MAX_RANKS];
	double4 *peer_x[ADMMB200_MAX_RANKS];
};

// Halo exchange without flags or fences ("flag in data", as in NCCL's LL protocol): every published
// component is ONE aligned 64-bit word {value, tag}, tag = tag_base | (pass + 1).  A 64-bit store is
// single-copy atomic, so a reader that sees the expected tag has the value of exactly that pass; it
// simply re-reads until then.  There is no back-pressure, so the buffer is double-buffered by sweep
// parity: a part can only overwrite a word two sweeps later, and to get there it must have received --
// from every neighbour -- values those neighbours computed after they consumed the word.
__device__ __forceinline__ void ll_store(uint2 *p, float v, unsigned int tag) {
	asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_store_sys(uint2 *p, float v, unsigned int tag) { // into a peer GPU's memory
	asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ll_load(const uint2 *p) {
	uint2 v;
	asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ float4 ldcg_f4(const float4 *p) { return __ldcg(p); }

template <int T>
__device__ __forceinline__ void res32_gather(const float *s_val, const uint16_t *s_col, const float4 *s_d, int r0, int r1, int lane,
	float &sx, float &sy, float &sz)
{
	sx = 0.f; sy = 0.f; sz = 0.f;
#pragma unroll 4
	for (int r = r0; r < r1; ++r) {
		const int c = s_col[r * 32 + lane];
		const float a = s_val[r * 32 + lane];
		const float4 dv = s_d[c];
		sx = fmaf(a, dv.x, sx); sy = fmaf(a, dv.y, sy); sz = fmaf(a, dv.z, sz);
	}
#pragma unroll
	for (int o = 1; o < T; o <<= 1) {
		sx += __shfl_xor_sync(0xffffffffu, sx, o);
		sy += __shfl_xor_sync(0xffffffffu, sy, o);
		sz += __shfl_xor_sync(0xffffffffu, sz, o);
	}
}

// PROF / DBG: cycle counters and timing experiments (tools/gs_prof.py) are a separate instantiation, so the
// production kernel carries none of their state (it used to cost 144 B of spills per thread).
template <int T, bool PROF>
__global__ void __launch_bounds__(ADMMB200_RES_THREADS, 1) mcgs_resident_f32_kernel(McgsRes32Params R)
{
	constexpr int G = 32 / T;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ double red[32];
	__shared__ __align__(8) uint64_t tma_bar;
	__shared__ int s_nbr[192];
	__shared__ int s_decision;
	const McgsParams &P = R.base;
	const long long t_kernel = PROF ? clock64() : 0;
	const PartDesc d = R.parts[R.part0 + blockIdx.x];
	const int tid = threadIdx.x, lane = tid & 31, sub = lane % T, grp = lane / T, warp = tid >> 5, n_warps = blockDim.x >> 5;
	const int n_loc = d.n_own + d.n_halo, C = P.n_colors;

	// shared-memory layout: must match ResidentPlan::layout(mode 1)
	size_t o = 0;
	auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
	float4 *s_d = (float4 *)(smem + take(16 * (size_t)n_loc));
	float *s_val = (float *)(smem + take(res32_val_region((size_t)d.n_rows, (size_t)n_loc)));
	uint16_t *s_col = (uint16_t *)(smem + take(sizeof(uint16_t) * 32 * (size_t)d.n_rows));
	int *s_gid = (int *)(smem + take(sizeof(int) * (size_t)n_loc));
	int *s_srow = (int *)(smem + take(sizeof(int) * ((size_t)d.n_slices + 1)));
	short *s_snode = (short *)(smem + take(sizeof(short) * (size_t)G * d.n_slices));
	int *s_cslice = (int *)(smem + take(sizeof(int) * (2 * (size_t)C + 1)));
	int *s_hcol = (int *)(smem + take(sizeof(int) * ((size_t)C + 1)));

	// ---- stage the part: matrix by TMA bulk copy, index tables by plain loads, d = 0 ----
	const uint32_t val_bytes = (uint32_t)(sizeof(float) * 32 * (size_t)d.n_rows), col_bytes = (uint32_t)(sizeof(uint16_t) * 32 * (size_t)d.n_rows);
	if (tid == 0) mbar_init(&tma_bar, 1);
	__syncthreads();
	if (tid == 0 && d.n_rows > 0) {
		mbar_expect_tx(&tma_bar, val_bytes + col_bytes);
		const unsigned char *gv = (const unsigned char *)(R.val + d.ent_off);
		const unsigned char *gc = (const unsigned char *)(R.col + d.ent_off);
		const uint32_t chunk = 32768;
		for (uint32_t at = 0; at < val_bytes; at += chunk) bulk_g2s((unsigned char *)s_val + at, gv + at, min(chunk, val_bytes - at), &tma_bar);
		for (uint32_t at = 0; at < col_bytes; at += chunk) bulk_g2s((unsigned char *)s_col + at, gc + at, min(chunk, col_bytes - at), &tma_bar);
	}
	for (int i = tid; i < n_loc; i += blockDim.x) { s_gid[i] = R.gid[d.gid_off + i]; s_d[i] = make_float4(0.f, 0.f, 0.f, 0.f); }
	for (int i = tid; i <= d.n_slices; i += blockDim.x) s_srow[i] = R.slice_row[d.slice_off + i];
	for (int i = tid; i < G * d.n_slices; i += blockDim.x) s_snode[i] = R.slice_node[d.snode_off + i];
	for (int i = tid; i <= 2 * C; i += blockDim.x) s_cslice[i] = R.color_slice[d.cslice_off + i];
	for (int i = tid; i <= C; i += blockDim.x) s_hcol[i] = R.halo_color[d.hcolor_off + i];
	for (int i = tid; i < d.n_nbr; i += blockDim.x) s_nbr[i] = R.nbr[d.nbr_off + i];
	if (d.n_rows > 0) mbar_wait(&tma_bar, 0);
	__syncthreads();

	const long long t_staged = PROF ? clock64() : 0;
	unsigned int bar_target = 0;
	const bool check = P.tol2 > 0.0;
	const float omega = (float)P.omega, one_m_omega = (float)(1.0 - P.omega), lb_scale = (float)(1.0 / P.omega - 1.0);
	float4 *nb = R.nodebuf + 2 * (size_t)d.own_off;

	// ---- r0 = b - A x_ref in fp64 (x_ref = P.x, read-only until the very end), and |b|^2 ----
	// With the EXACT matrix: the rows of L sum to zero (translation invariance), rounding its entries to
	// fp32 breaks that by 6e-8 |a_ii| per row, which multiplied by |x_ref| ~ metres would be a visible
	// force; multiplied by the millimetre increment in the sweeps it is not.
	{
		double b2 = 0;
		const double *g_val64 = R.val64 + d.ent_off;
		for (int sl = warp; sl < d.n_slices; sl += n_warps) {
			const int l = s_snode[sl * G + grp];
			double sx = 0, sy = 0, sz = 0;
			for (int r = s_srow[sl]; r < s_srow[sl + 1]; ++r) {
				const int c = s_col[r * 32 + lane];
				const double a = __ldg(&g_val64[(size_t)r * 32 + lane]);
				const double4 xc = ld_node(&P.x[s_gid[c]]);
				sx += a * xc.x; sy += a * xc.y; sz += a * xc.z;
			}
#pragma unroll
			for (int o2 = 1; o2 < T; o2 <<= 1) {
				sx += __shfl_xor_sync(0xffffffffu, sx, o2);
				sy += __shfl_xor_sync(0xffffffffu, sy, o2);
				sz += __shfl_xor_sync(0xffffffffu, sz, o2);
			}
			if (sub == 0 && l >= 0) {
				const int node = s_gid[l];
				const double4 bi = ld_node(&P.b[node]), xi = ld_node(&P.x[node]);
				const double a0 = __ldg(&P.diag[3 * node]), a1 = __ldg(&P.diag[3 * node + 1]), a2 = __ldg(&P.diag[3 * node + 2]);
				const int ps = P.has_pins ? __ldg(&P.pin_slot[node]) : -1;
				nb[2 * l] = make_float4((float)(bi.x - sx - a0 * xi.x), (float)(bi.y - sy - a1 * xi.y), (float)(bi.z - sz - a2 * xi.z), __int_as_float(ps + 1));
				const unsigned int dm = R.dest_mask ? __ldg(&R.dest_mask[node]) : 0u;
				nb[2 * l + 1] = make_float4((float)(1.0 / a0), (float)(1.0 / a1), (float)(1.0 / a2), __uint_as_float(dm));
				b2 += bi.x * bi.x + bi.y * bi.y + bi.z * bi.z;
			}
		}
		if (check) {
			double s = block_sum(b2, red); // b_norm = |b|^2 (src/NodalMultiColorGS.hpp:92)
			if (tid == 0) atomicAdd(&P.resid[0], s);
			grid_barrier(P.barrier, bar_target, gridDim.x); // the only grid-wide barrier of a solve
		} else __syncthreads();
	}
	const double thresh = check ? 4.0 * P.tol2 * __ldcg(&P.resid[0]) : 0.0;

	double lb = 0;
	unsigned int pass_tag = 0;   // tag of the pass being computed
	size_t pub_off = 0;           // offset of the current sweep parity's buffer in dglob
	auto do_slice = [&](int sl, bool to_global, bool last) {
		const int l = s_snode[sl * G + grp];
		const bool owner = (sub == 0 && l >= 0);
		float4 rb = make_float4(0.f, 0.f, 0.f, 0.f), ia = rb;
		if (owner) { rb = nb[2 * l]; ia = nb[2 * l + 1]; }
		float sx, sy, sz;
		res32_gather<T>(s_val, s_col, s_d, s_srow[sl], s_srow[sl + 1], lane, sx, sy, sz);
		if (owner) {
			const float4 dold = s_d[l];
			float4 dn;
			const int ps = __float_as_int(rb.w) - 1;
			if (ps >= 0) {
				// pinned node (src/NodalMultiColorGS.hpp:111-117): x = pin, i.e. d = pin - x_ref
				const double4 xr = ld_node(&P.x[s_gid[l]]);
				dn = make_float4((float)(P.pin_pos[3 * ps] - xr.x), (float)(P.pin_pos[3 * ps + 1] - xr.y), (float)(P.pin_pos[3 * ps + 2] - xr.z), 0.f);
			} else {
				// segment_update (src/NodalMultiColorGS.hpp:180-215) on the increment
				const float g0 = (rb.x - sx) * ia.x, g1 = (rb.y - sy) * ia.y, g2 = (rb.z - sz) * ia.z;
				dn = make_float4(fmaf(omega, g0, one_m_omega * dold.x), fmaf(omega, g1, one_m_omega * dold.y), fmaf(omega, g2, one_m_omega * dold.z), 0.f);
				bool hit = false;
				if (P.n_obstacles > 0) {
					// the obstacle test needs absolute positions
					const double4 xr = ld_node(&P.x[s_gid[l]]);
					double gs[3] = {xr.x + (double)g0, xr.y + (double)g1, xr.z + (double)g2};
					double nx[3] = {xr.x + (double)dn.x, xr.y + (double)dn.y, xr.z + (double)dn.z};
					hit = mcgs_collide(P.obs, P.n_obstacles, gs, nx);
					if (hit) dn = make_float4((float)(nx[0] - xr.x), (float)(nx[1] - xr.y), (float)(nx[2] - xr.z), 0.f);
				}
				if (last && !hit) {
					const float rx = lb_scale * (dn.x - dold.x) / ia.x, ry = lb_scale * (dn.y - dold.y) / ia.y, rz = lb_scale * (dn.z - dold.z) / ia.z;
					lb += (double)rx * rx + (double)ry * ry + (double)rz * rz;
				}
			}
			s_d[l] = dn;
			if (to_global) {
				const size_t at = pub_off + 3 * (size_t)s_gid[l];
				uint2 *w = R.dglob + at;
				ll_store(w, dn.x, pass_tag); ll_store(w + 1, dn.y, pass_tag); ll_store(w + 2, dn.z, pass_tag);
				unsigned int dm = __float_as_uint(ia.w);
				while (dm) { // peers that read this node
					const int q = __ffs(dm) - 1; dm &= dm - 1;
					uint2 *wq = R.peer_dglob[q] + at;
					ll_store_sys(wq, dn.x, pass_tag); ll_store_sys(wq + 1, dn.y, pass_tag); ll_store_sys(wq + 2, dn.z, pass_tag);
				}
			}
		}
	};

	// Warp roles (see mcgs_resident.cuh), split in proportion to the boundary / interior work of this part.
	int n_bwarps;
	{
		int mb = 0, mi = 0;
		for (int c = 0; c < C; ++c) { mi = max(mi, s_cslice[2 * c + 1] - s_cslice[2 * c]); mb = max(mb, s_cslice[2 * c + 2] - s_cslice[2 * c + 1]); }
		n_bwarps = (mb + mi) > 0 ? (n_warps * mb + (mb + mi) / 2) / (mb + mi) : n_warps / 2;
		n_bwarps = min(max(n_bwarps, 1), n_warps - 1);
	}
	const int n_iwarps = n_warps - n_bwarps, n_bthreads = 32 * n_bwarps;
	const bool bwarp = warp < n_bwarps;
	const size_t buf_stride = 3 * (size_t)R.n_nodes_total;

	// pulls the halo values of colour `cp` published with tag `tag` in buffer `buf` into shared memory
	auto refresh = [&](int cp, const uint2 *buf, unsigned int tag, int t0, int nt) {
		for (int h = s_hcol[cp] + t0; h < s_hcol[cp + 1]; h += nt) {
			const uint2 *w = buf + 3 * (size_t)s_gid[d.n_own + h];
			uint2 a, b, c;
			do { a = ll_load(w); b = ll_load(w + 1); c = ll_load(w + 2); } while ((a.y != tag || b.y != tag || c.y != tag) && !(PROF && (R.dbg & 1)));
			s_d[d.n_own + h] = make_float4(__uint_as_float(a.x), __uint_as_float(b.x), __uint_as_float(c.x), 0.f);
		}
	};

	int it = 0;
	unsigned int pass = 0; // passes done so far
	long long pw = 0, pc = 0, pp = 0, pi = 0, pk = 0;
	const long long t_begin = PROF ? clock64() : 0;
	for (; it < P.iters; ++it) {
		lb = 0;
		for (int color = 0; color < C; ++color) {
			const int s0 = s_cslice[2 * color], s1 = s_cslice[2 * color + 1], s2 = s_cslice[2 * color + 2];
			const bool last = check && (color == C - 1);
			pass_tag = R.tag_base | (pass + 1);
			pub_off = (size_t)(it & 1) * buf_stride;
			long long t0 = 0, t1 = 0, t2 = 0;
			if (PROF) t0 = clock64();
			if (bwarp) {
				if (pass > 0 && !(PROF && (R.dbg & 2))) {
					// what the neighbours changed in the previous pass: the halo nodes of that pass's colour
					const int cp = (color + C - 1) % C;
					const int it_prev = color > 0 ? it : it - 1;
					refresh(cp, R.dglob + (size_t)(it_prev & 1) * buf_stride, R.tag_base | pass, tid, n_bthreads);
					named_sync(1, n_bthreads);
				}
				if (PROF) t1 = clock64();
				if (!(PROF && (R.dbg & 8))) for (int sl = s1 + warp; sl < s2; sl += n_bwarps) do_slice(sl, true, last);
				if (PROF) t2 = clock64();
			} else {
				if (!(PROF && (R.dbg & 4))) for (int sl = s0 + (warp - n_bwarps); sl < s1; sl += n_iwarps) do_slice(sl, false, last);
				if (PROF) t1 = clock64();
			}
			__syncthreads();
			++pass;
			if (PROF && tid == 0) { long long t3 = clock64(); pw += t1 - t0; pc += t2 - t1; pp += t3 - t2; }
			if (PROF && tid == n_bthreads) pi += t1 - t0;
		}
		const long long tk0 = PROF ? clock64() : 0;
		if (check) {
			// "converged?" (see mcgs_resident.cuh).  A part whose own rows prove |b - A x|^2 >= 4 tol^2 |b|^2
			// knows the answer without talking to anyone; it only leaves a note for parts that cannot.
			double s = block_sum(lb, red);
			if (tid == 0) {
				int decision = -1;
				if (s >= thresh) { atomicAdd(&R.sweep_arrive[it], 0x10001u); decision = 1; } // arrived + proved, in one word
				else {
					if (s > 0.0) atomicAdd(&P.resid_lb[it], s);
					__threadfence();
					atomicAdd(&R.sweep_arrive[it], 1u);
					while (decision < 0) {
						const unsigned int v = ld_relaxed_u32(&R.sweep_arrive[it]);
						if ((v >> 16) != 0u) decision = 1;             // somebody proved "not converged"
						else if ((v & 0xffffu) == gridDim.x) {          // everybody is here and nobody could: look at the summed bound
							fence_acq_rel_gpu();
							decision = (__ldcg(&P.resid_lb[it]) >= thresh) ? 1 : 0;
						}
					}
				}
				s_decision = decision;
			}
			__syncthreads();
			const bool proven_unconverged = s_decision == 1;
			__syncthreads();
			if (!proven_unconverged) {
				const double b2 = __ldcg(&P.resid[0]);
				// exact residual b - A x = r0 - A d (src/NodalMultiColorGS.hpp:136-139).  Every part takes this
				// branch; the last colour's halo values are the only ones not pulled in yet.
				refresh(C - 1, R.dglob + (size_t)(it & 1) * buf_stride, R.tag_base | pass, tid, (int)blockDim.x);
				__syncthreads();
				double acc = 0;
				for (int sl = warp; sl < d.n_slices; sl += n_warps) {
					const int l = s_snode[sl * G + grp];
					float sx, sy, sz;
					res32_gather<T>(s_val, s_col, s_d, s_srow[sl], s_srow[sl + 1], lane, sx, sy, sz);
					if (sub == 0 && l >= 0) {
						const float4 rb = nb[2 * l], ia = nb[2 * l + 1], dv = s_d[l];
						double rx = (double)rb.x - (double)sx - (double)dv.x / (double)ia.x;
						double ry = (double)rb.y - (double)sy - (double)dv.y / (double)ia.y;
						double rz = (double)rb.z - (double)sz - (double)dv.z / (double)ia.z;
						acc += rx * rx + ry * ry + rz * rz;
					}
				}
				double sres = block_sum(acc, red);
				if (tid == 0) atomicAdd(&P.resid[1 + it], sres);
				grid_barrier(P.barrier, bar_target, gridDim.x);
				double r2 = __ldcg(&P.resid[1 + it]);
				if (r2 / b2 < P.tol2) break;
			}
		}
		if (PROF) pk += clock64() - tk0;
	}
	// x = x_ref + d: the only write to the positions
	const long long t_loop_end = PROF ? clock64() : 0;
	__syncthreads();
	for (int l = tid; l < d.n_own; l += blockDim.x) {
		const int node = s_gid[l];
		const double4 xr = P.x[node];
		const float4 dv = s_d[l];
		const double nx0 = xr.x + (double)dv.x, nx1 = xr.y + (double)dv.y, nx2 = xr.z + (double)dv.z;
		st_node(&P.x[node], nx0, nx1, nx2);
		if (R.world > 1) { // ghost copies on the peers (their next local step and r0 read them)
			unsigned int dm = __float_as_uint(nb[2 * l + 1].w);
			while (dm) { const int q = __ffs(dm) - 1; dm &= dm - 1; st_node(&R.peer_x[q][node], nx0, nx1, nx2); }
		}
	}
	if (blockIdx.x == 0 && tid == 0) *P.iters_done = it;
	if (PROF && tid == 0) {
		unsigned long long *q = R.prof + 16 * blockIdx.x;
		q[0] = (unsigned long long)pw; q[1] = (unsigned long long)pc; q[2] = (unsigned long long)pp; q[3] = (unsigned long long)(clock64() - t_kernel);
		q[13] = (unsigned long long)(t_staged - t_kernel); q[14] = (unsigned long long)(t_begin - t_staged); q[15] = (unsigned long long)(t_loop_end - t_begin);
	}
	if (PROF && tid == n_bthreads) R.prof[16 * blockIdx.x + 4] = (unsigned long long)pi;
	if (PROF && tid == 0) R.prof[16 * blockIdx.x + 5] = (unsigned long long)pk;
	(void)pk; (void)pw; (void)pc; (void)pp; (void)pi;
}

} // namespace admmb200
