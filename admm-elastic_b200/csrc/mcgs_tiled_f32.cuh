// mcgs_tiled_f32.cuh -- NodalMultiColorGS::solve (src/NodalMultiColorGS.hpp:60-146), the production global step since
// round 2: the shared-memory-resident fp32 sweep on the increment (numerics: mcgs_resident_f32.cuh -- SOR on A d = r0
// around the fp64 anchor x_ref) cut into SHORT tasks.
//
// What round 1's static-ownership kernel (mcgs_owned_f32.cuh) measured on the 1M-tet beam: a colour pass costs ~5 900
// cycles, of which the shared-memory pipe needs ~1 800; the rest is a chain of latencies.  One warp gathers a slice of 32
// nodes with one lane per node: ~18 dependent row-steps, ~80 cycles each when the warp runs alone
// (tools/micro/gather_bench.cu: 1 450 cycles per slice with 1 active warp, 285 with 16) -- and in a pass only 4-5 warps
// hold a boundary slice, which is exactly the work the neighbours wait for.  A 100k-tet mesh (175 nodes per part) still
// needs 4 460 cycles per pass: the pass time is latency, not work.
//
// Here a task is 8 nodes x 4 lanes (T = 4): a lane walks ceil(len / 4) ~ 5 row-steps, all loads of a task in flight at
// once, two shuffle rounds add the partial sums.  A pass of colour c is
//     interior tasks (need no halo)            all 16 warps, round-robin
//     poll the mailboxes of colour c-1         every thread one slot; the values arrived while the interior ran
//     CTA barrier
//     boundary tasks + publish                 all 16 warps, ~20 tasks
//     CTA barrier (carries the "converged?" vote in the last colour)
// so the neighbours' values are published ~400 cycles after they became computable instead of ~2 500, and nothing a
// task needs comes from L2: the matrix is pre-scaled on the host (a'_ij = omega a_ij / a_ii, so the update is
// d_new = (1 - omega) d_old + (rbs - sum a'_ij d_j) with rbs = omega r0 / a_ii), rbs lives in shared memory, a_ii rides
// in the unused fourth component of the node's float4 increment (negative: pinned node), and the first two mailbox
// slots of a boundary node come with one 16-byte load issued before the gather.  No per-slice state in registers, no
// unrolled per-slice code.
//
// Needs: equal x/y/z masses per node (a_ii is one number), no obstacles (they use mcgs_owned_f32.cuh), fp32 elements.
// Halo protocol, convergence test, pins, multi-GPU pushes: unchanged from mcgs_owned_f32.cuh (same mailboxes, same
// {value, tag} words), so the two kernels are interchangeable behind launch_mcgs_resident.
// Shared-memory layout: ResidentPlan::layout(mode 2).
#pragma once
#include "mcgs_owned_f32.cuh"

namespace admmb200 {

#define ADMMB200_TILED_THREADS 512

struct McgsTiledExtra {
	const float *val_scaled;  // omega a_ij / a_ii, same indexing as col
	const uint4 *dest4;       // [n_nodes] by (own_off + local id): {slot0, slot1, first entry in dest_slot, readers}
};

template <int T, bool PROF>
__global__ void __launch_bounds__(ADMMB200_TILED_THREADS, 1) mcgs_tiled_f32_kernel(McgsRes32Params R, McgsTiledExtra X)
{
	constexpr int NT = ADMMB200_TILED_THREADS, NW = NT / 32, G = 32 / T;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ double red[32];
	__shared__ __align__(8) uint64_t tma_bar;
	__shared__ int s_decision;
	const McgsParams &P = R.base;
	const long long t_kernel = PROF ? clock64() : 0;
	const PartDesc d = R.parts[R.part0 + blockIdx.x];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = lane % T, grp = lane / T;
	const int n_loc = d.n_own + d.n_halo, C = P.n_colors;

	// shared-memory layout: ResidentPlan::layout(mode 2)
	size_t o = 0;
	auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
	float4 *s_d = (float4 *)(smem + take(16 * (size_t)n_loc));
	float *s_val = (float *)(smem + take(res32_val_region((size_t)d.n_rows, (size_t)n_loc)));
	double *s_x = (double *)s_val; // x_ref of owned + halo nodes until r0 is formed; the matrix values arrive afterwards
	uint16_t *s_col = (uint16_t *)(smem + take(sizeof(uint16_t) * 32 * (size_t)d.n_rows));
	float *s_rb = (float *)(smem + take(sizeof(float) * 3 * (size_t)d.n_own));
	int *s_srow = (int *)(smem + take(sizeof(int) * ((size_t)d.n_slices + 1)));
	int *s_sinfo = (int *)(smem + take(sizeof(int) * (size_t)d.n_slices)); // first local node | valid nodes << 16
	int *s_cslice = (int *)(smem + take(sizeof(int) * (2 * (size_t)C + 1)));
	int *s_hcol = (int *)(smem + take(sizeof(int) * ((size_t)C + 1)));

	// ---- stage the part: column indices by TMA bulk copy, tables by plain loads, d = 0, x_ref into the value region ----
	const uint32_t val_bytes = (uint32_t)(sizeof(float) * 32 * (size_t)d.n_rows), col_bytes = (uint32_t)(sizeof(uint16_t) * 32 * (size_t)d.n_rows);
	if (tid == 0) mbar_init(&tma_bar, 1);
	__syncthreads();
	if (tid == 0 && d.n_rows > 0) {
		mbar_expect_tx(&tma_bar, col_bytes);
		const unsigned char *gc = (const unsigned char *)(R.col + d.ent_off);
		const uint32_t chunk = 32768;
		for (uint32_t at = 0; at < col_bytes; at += chunk) bulk_g2s((unsigned char *)s_col + at, gc + at, min(chunk, col_bytes - at), &tma_bar);
	}
	for (int i = tid; i < n_loc; i += NT) {
		const int g = __ldg(&R.gid[d.gid_off + i]);
		s_d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		const double4 xg = ld_node_cg(&P.x[g]); // L2-coherent: peers and other parts wrote these at the end of the previous solve
		s_x[3 * i] = xg.x; s_x[3 * i + 1] = xg.y; s_x[3 * i + 2] = xg.z;
	}
	for (int i = tid; i <= d.n_slices; i += NT) s_srow[i] = __ldg(&R.slice_row[d.slice_off + i]);
	for (int sl = tid; sl < d.n_slices; sl += NT) {
		int first = -1, cnt = 0;
		for (int g = 0; g < G; ++g) { const int l = (int)__ldg(&R.slice_node[d.snode_off + sl * G + g]); if (l >= 0) { if (first < 0) first = l; ++cnt; } }
		s_sinfo[sl] = (first < 0 ? 0 : first) | (cnt << 16);
	}
	for (int i = tid; i <= 2 * C; i += NT) s_cslice[i] = __ldg(&R.color_slice[d.cslice_off + i]);
	for (int i = tid; i <= C; i += NT) s_hcol[i] = __ldg(&R.halo_color[d.hcolor_off + i]);
	if (d.n_rows > 0) mbar_wait(&tma_bar, 0);
	__syncthreads();

	const long long t_staged = PROF ? clock64() : 0;
	unsigned int bar_target = 0;
	const bool check = P.tol2 > 0.0;
	const float omega = (float)P.omega, one_m_omega = (float)(1.0 - P.omega), inv_omega = (float)(1.0 / P.omega), lb_scale = (float)(1.0 / P.omega - 1.0);
	float4 *pinbuf = R.nodebuf + 2 * (size_t)d.own_off; // [l]: pin - x_ref of a pinned node (read in the first sweep only)

	// ---- r0 = b - A x_ref in fp64 with the EXACT matrix (see mcgs_resident_f32.cuh), and |b|^2.  The fp64 values stream
	// from global memory once (coalesced); x_ref is gathered from shared memory.  rbs = omega r0 / a_ii -> shared memory. ----
	{
		double b2 = 0;
		const double *g_val64 = R.val64 + d.ent_off;
		for (int sl = warp; sl < d.n_slices; sl += NW) {
			const int info = s_sinfo[sl], l = (info & 0xffff) + grp;
			const bool valid = grp < (info >> 16);
			double sx = 0, sy = 0, sz = 0;
			const int r0 = s_srow[sl], r1 = s_srow[sl + 1];
			for (int r = r0; r < r1; ++r) {
				const double a = __ldg(g_val64 + (size_t)r * 32 + lane);
				const double *xc = s_x + 3 * (int)s_col[r * 32 + lane];
				sx += a * xc[0]; sy += a * xc[1]; sz += a * xc[2];
			}
#pragma unroll
			for (int o2 = 1; o2 < T; o2 <<= 1) {
				sx += __shfl_xor_sync(0xffffffffu, sx, o2);
				sy += __shfl_xor_sync(0xffffffffu, sy, o2);
				sz += __shfl_xor_sync(0xffffffffu, sz, o2);
			}
			if (sub == 0 && valid) {
				const int node = __ldg(&R.gid[d.gid_off + l]);
				const double4 bi = ld_node(&P.b[node]);
				const double xi0 = s_x[3 * l], xi1 = s_x[3 * l + 1], xi2 = s_x[3 * l + 2];
				const double a0 = __ldg(&P.diag[3 * node]); // equal x/y/z masses: one diagonal value per node (checked on the host)
				const int ps = P.has_pins ? __ldg(&P.pin_slot[node]) : -1;
				const double sc = P.omega / a0;
				s_rb[3 * l] = (float)(sc * (bi.x - sx - a0 * xi0)); s_rb[3 * l + 1] = (float)(sc * (bi.y - sy - a0 * xi1)); s_rb[3 * l + 2] = (float)(sc * (bi.z - sz - a0 * xi2));
				float aw = (float)a0;
				if (ps >= 0) {
					// pinned node (src/NodalMultiColorGS.hpp:111-117): x = pin, i.e. d = pin - x_ref from its first update on
					aw = -aw;
					pinbuf[l] = make_float4((float)(P.pin_pos[3 * ps] - xi0), (float)(P.pin_pos[3 * ps + 1] - xi1), (float)(P.pin_pos[3 * ps + 2] - xi2), 0.f);
				}
				s_d[l].w = aw;
				b2 += bi.x * bi.x + bi.y * bi.y + bi.z * bi.z;
			}
		}
		__syncthreads(); // nobody reads x_ref from the value region any more: the fp32 values may land there
		if (tid == 0 && d.n_rows > 0) {
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			mbar_expect_tx(&tma_bar, val_bytes);
			const unsigned char *gv = (const unsigned char *)(X.val_scaled + d.ent_off);
			const uint32_t chunk = 32768;
			for (uint32_t at = 0; at < val_bytes; at += chunk) bulk_g2s((unsigned char *)s_val + at, gv + at, min(chunk, val_bytes - at), &tma_bar);
		}
		if (check) {
			double s = block_sum(b2, red); // b_norm = |b|^2 (src/NodalMultiColorGS.hpp:92)
			if (tid == 0) atomicAdd(&P.resid[0], s);
			grid_barrier(P.barrier, bar_target, gridDim.x); // the only grid-wide barrier of a solve
		}
		if (d.n_rows > 0) mbar_wait(&tma_bar, 1);
		__syncthreads();
	}
	const double thresh = check ? 4.0 * P.tol2 * __ldcg(&P.resid[0]) : 0.0;
	const size_t TS = (size_t)R.total_slots, buf_stride = 3 * TS; // dglob: [sweep parity][x | y | z][slot]
	const uint4 *dest4 = X.dest4 + d.own_off;
	long long p_int = 0, p_poll = 0, p_bnd = 0, p_bar = 0, n_spin = 0;

	// One task: the G nodes of slice sl, T lanes each.  Returns this lane's contribution to the residual bound.
	auto do_task = [&](int sl, bool bnd, bool last, int it, unsigned int pass_tag, size_t pub_off) -> float {
		const int info = s_sinfo[sl], l = (info & 0xffff) + grp;
		const bool owner = sub == 0 && grp < (info >> 16);
		uint4 dst = make_uint4(0u, 0u, 0u, 0u);
		if (bnd && owner) dst = __ldg(&dest4[l]); // where this node is published: in flight during the gather
		const int r0 = s_srow[sl], r1 = s_srow[sl + 1];
		float sx = 0.f, sy = 0.f, sz = 0.f;
		const float *v = s_val + r0 * 32 + lane;
		const uint16_t *c = s_col + r0 * 32 + lane;
		const int n = r1 - r0;
		for (int r = 0; r < n; r += 6) {
			int cc[6]; float a[6];
#pragma unroll
			for (int j = 0; j < 6; ++j) {
				const int rr = min(r + j, n - 1);
				cc[j] = c[rr * 32];
				a[j] = (r + j < n) ? v[rr * 32] : 0.f;
			}
#pragma unroll
			for (int j = 0; j < 6; ++j) {
				const float4 dv = s_d[cc[j]];
				sx = fmaf(a[j], dv.x, sx); sy = fmaf(a[j], dv.y, sy); sz = fmaf(a[j], dv.z, sz);
			}
		}
#pragma unroll
		for (int o2 = 1; o2 < T; o2 <<= 1) {
			sx += __shfl_xor_sync(0xffffffffu, sx, o2);
			sy += __shfl_xor_sync(0xffffffffu, sy, o2);
			sz += __shfl_xor_sync(0xffffffffu, sz, o2);
		}
		float lbv = 0.f;
		if (owner) {
			const float4 dold = s_d[l];
			float4 dn;
			if (dold.w < 0.f) {
				dn = dold;
				if (it == 0) { const float4 pv = pinbuf[l]; dn = make_float4(pv.x, pv.y, pv.z, dold.w); }
			} else {
				// segment_update (src/NodalMultiColorGS.hpp:180-215) on the increment, matrix and right-hand side pre-scaled by omega / a_ii
				const float rb0 = s_rb[3 * l], rb1 = s_rb[3 * l + 1], rb2 = s_rb[3 * l + 2];
				dn = make_float4(fmaf(one_m_omega, dold.x, rb0 - sx), fmaf(one_m_omega, dold.y, rb1 - sy), fmaf(one_m_omega, dold.z, rb2 - sz), dold.w);
				if (last) {
					// residual row right after the update: a_ii (1/omega - 1) (d_new - d_old); steering only (4x margin)
					const float rx = lb_scale * dold.w * (dn.x - dold.x), ry = lb_scale * dold.w * (dn.y - dold.y), rz = lb_scale * dold.w * (dn.z - dold.z);
					lbv = rx * rx + ry * ry + rz * rz;
				}
			}
			s_d[l] = dn;
			if (bnd && !(PROF && (R.dbg & 16))) {
				const int cnt = (int)dst.w;
				auto put = [&](unsigned int ent) {
					const unsigned int q = ent >> 27;
					const size_t at = pub_off + (size_t)(ent & 0x7ffffffu);
					if ((int)q == R.rank) { uint2 *w = R.dglob + at; ll_store(w, dn.x, pass_tag); ll_store(w + TS, dn.y, pass_tag); ll_store(w + 2 * TS, dn.z, pass_tag); }
					else { uint2 *w = R.peer_dglob[q] + at; ll_store_sys(w, dn.x, pass_tag); ll_store_sys(w + TS, dn.y, pass_tag); ll_store_sys(w + 2 * TS, dn.z, pass_tag); }
				};
				if (cnt > 0) put(dst.x);
				if (cnt > 1) put(dst.y);
				for (int e = (int)dst.z + 2; e < (int)dst.z + cnt; ++e) put(__ldg(&R.dest_slot[e])); // a corner node read by more than two parts
			}
		}
		(void)inv_omega; (void)omega;
		return lbv;
	};

	// Pulls the halo values of colour `cp` published with tag `tag` in buffer `buf` into shared memory.
	auto refresh = [&](int cp, const uint2 *buf, unsigned int tag) {
		const int end = s_hcol[cp + 1];
		for (int h = s_hcol[cp] + tid; h < end; h += NT) {
			const uint2 *w = buf + (size_t)d.slot_off + h; // consecutive threads, consecutive slots: coalesced polls
			uint2 a, b, c;
			a = ll_load(w); b = ll_load(w + TS); c = ll_load(w + 2 * TS);
			while (a.y != tag || b.y != tag || c.y != tag) { if (PROF) ++n_spin; a = ll_load(w); b = ll_load(w + TS); c = ll_load(w + 2 * TS); }
			s_d[d.n_own + h] = make_float4(__uint_as_float(a.x), __uint_as_float(b.x), __uint_as_float(c.x), 0.f);
		}
	};

	int it = 0;
	bool converged = false;
	unsigned int pass = 0; // passes done so far
	const long long t_begin = PROF ? clock64() : 0;
	for (; it < P.iters; ++it) {
		float lb = 0.f; // this lane's part of the lower bound of |b - A x|^2 (rows of the last colour)
		const size_t pub_off = (size_t)(it & 1) * buf_stride;
		for (int color = 0; color < C; ++color) {
			const bool last = check && (color == C - 1);
			const unsigned int pass_tag = R.tag_base | (pass + 1);
			const int i0 = s_cslice[2 * color], b0 = s_cslice[2 * color + 1], b1 = s_cslice[2 * color + 2];
			long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
			if (PROF) c0 = clk_ordered();
			// interior tasks: nobody outside this part reads them, and they read no halo value that is still in flight
			for (int sl = i0 + warp; sl < b0; sl += NW) lb += do_task(sl, false, last, it, pass_tag, pub_off);
			if (PROF) c1 = clk_ordered();
			if (pass > 0 && !(PROF && (R.dbg & 2))) {
				// what the neighbours changed in the previous pass: the halo nodes of that pass's colour
				const int cp = (color + C - 1) % C;
				const int it_prev = color > 0 ? it : it - 1;
				refresh(cp, R.dglob + (size_t)(it_prev & 1) * buf_stride, R.tag_base | pass);
			}
			__syncthreads(); // halo values of colour c-1 and the interior values of colour c are in place
			if (PROF) c2 = clk_ordered();
			// boundary tasks, spread over the warps starting where the interior tasks ended
			{
				const int shift = (b0 - i0) % NW;
				for (int sl = b0 + ((warp + NW - shift) % NW); sl < b1; sl += NW) lb += do_task(sl, true, last, it, pass_tag, pub_off);
			}
			if (PROF) c3 = clk_ordered();
			++pass;
			if (!last) {
				__syncthreads();
				if (PROF) { p_int += c1 - c0; p_poll += c2 - c1; p_bnd += c3 - c2; p_bar += clk_ordered() - c3; }
				continue;
			}

			// ---- "converged?" after the sweep (see kernels.cuh).  The barrier that ends the pass also answers
			// "can any lane of this part prove |b - A x|^2 >= 4 tol^2 |b|^2 from its own rows alone?" ----
			const int proven = __syncthreads_or((double)lb >= thresh && lb > 0.f);
			if (PROF) { p_int += c1 - c0; p_poll += c2 - c1; p_bnd += c3 - c2; p_bar += clk_ordered() - c3; }
			if (proven) {
				// ONE atomic carries both "arrived" (low 16 bits) and "proved it" (high bits), see mcgs_owned_f32.cuh
				if (tid == 0) atomicAdd(&R.sweep_arrive[it], 0x10001u);
				continue;
			}
			double s = block_sum((double)lb, red);
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
			if (proven_unconverged) continue;
			{
				const double b2 = __ldcg(&P.resid[0]);
				// exact residual b - A x = r0 - A d (src/NodalMultiColorGS.hpp:136-139).  Every part takes this
				// branch; the last colour's halo values are the only ones not pulled in yet.
				refresh(C - 1, R.dglob + pub_off, R.tag_base | pass);
				__syncthreads();
				double acc = 0;
				for (int sl = warp; sl < d.n_slices; sl += NW) {
					const int info = s_sinfo[sl], l = (info & 0xffff) + grp;
					float sx = 0.f, sy = 0.f, sz = 0.f;
					for (int r = s_srow[sl]; r < s_srow[sl + 1]; ++r) {
						const float a = s_val[r * 32 + lane];
						const float4 dv = s_d[s_col[r * 32 + lane]];
						sx = fmaf(a, dv.x, sx); sy = fmaf(a, dv.y, sy); sz = fmaf(a, dv.z, sz);
					}
#pragma unroll
					for (int o2 = 1; o2 < T; o2 <<= 1) {
						sx += __shfl_xor_sync(0xffffffffu, sx, o2);
						sy += __shfl_xor_sync(0xffffffffu, sy, o2);
						sz += __shfl_xor_sync(0xffffffffu, sz, o2);
					}
					if (sub == 0 && grp < (info >> 16)) {
						// r_i = r0_i - sum a_ij d_j - a_ii d_i = a_ii [ (rbs_i - s'_i) / omega - d_i ]
						const float4 dv = s_d[l];
						const double aii = fabs((double)dv.w);
						const double rx = aii * (((double)s_rb[3 * l] - (double)sx) / P.omega - (double)dv.x);
						const double ry = aii * (((double)s_rb[3 * l + 1] - (double)sy) / P.omega - (double)dv.y);
						const double rz = aii * (((double)s_rb[3 * l + 2] - (double)sz) / P.omega - (double)dv.z);
						acc += rx * rx + ry * ry + rz * rz;
					}
				}
				double sres = block_sum(acc, red);
				if (tid == 0) atomicAdd(&P.resid[1 + it], sres);
				grid_barrier(P.barrier, bar_target, gridDim.x);
				double r2 = __ldcg(&P.resid[1 + it]);
				if (r2 / b2 < P.tol2) converged = true; // last colour: the colour loop ends here anyway
			}
		}
		if (converged) break; // `it` stays the index of the sweep that converged, as in the reference
	}
	// x = x_ref + d: the only write to the positions
	const long long t_loop_end = PROF ? clock64() : 0;
	__syncthreads();
	for (int l = tid; l < d.n_own; l += NT) {
		const int node = __ldg(&R.gid[d.gid_off + l]);
		const double4 xr0 = P.x[node];
		const float4 dv = s_d[l];
		const double nx0 = xr0.x + (double)dv.x, nx1 = xr0.y + (double)dv.y, nx2 = xr0.z + (double)dv.z;
		st_node(&P.x[node], nx0, nx1, nx2);
		if (R.world > 1) { // ghost copies on the peers (their next local step and r0 read them)
			unsigned int dm = __ldg(&R.dest_mask[node]);
			while (dm) { const int q = __ffs(dm) - 1; dm &= dm - 1; st_node(&R.peer_x[q][node], nx0, nx1, nx2); }
		}
	}
	if (blockIdx.x == 0 && tid == 0) *P.iters_done = it;
	if (PROF && tid == 0) {
		unsigned long long *q = R.prof + 16 * blockIdx.x;
		q[0] = (unsigned long long)p_poll; q[1] = (unsigned long long)p_bnd; q[2] = (unsigned long long)p_bar; q[3] = (unsigned long long)(clock64() - t_kernel);
		q[5] = (unsigned long long)p_int; q[8] = (unsigned long long)n_spin;
		q[13] = (unsigned long long)(t_staged - t_kernel); q[14] = (unsigned long long)(t_begin - t_staged); q[15] = (unsigned long long)(t_loop_end - t_begin);
	}
	(void)p_int; (void)p_poll; (void)p_bnd; (void)p_bar; (void)n_spin;
}

} // namespace admmb200
