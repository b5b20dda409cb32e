// mcgs_owned_f32.cuh -- NodalMultiColorGS::solve (src/NodalMultiColorGS.hpp:60-146), the production
// global step: the shared-memory-resident fp32 sweep on the increment (see mcgs_resident_f32.cuh for
// the numerics: SOR on A d = r0 around the fp64 anchor x_ref) with STATIC OWNERSHIP of the work.
//
// mcgs_resident_f32_kernel walks slice tables every pass: slice -> node -> {r0, 1/a} from L2, row range
// from shared memory, and spills (80 registers at 768 threads).  Measured on the 1M-tet beam it spends
// 6 200 cycles per colour pass even with the halo exchange switched off -- for ~12 slices of work.  A
// pass is a latency chain, so this kernel removes every load from that chain that does not have to be
// there:
//   * a warp owns the same <= KMAX slices (32 nodes each) for the whole solve; lane = node.  r0, 1/a,
//     the node's ids, its row range and its role (boundary / interior / pinned) stay in REGISTERS from
//     the set-up to the last sweep.  A pass is: [poll halo] -> gather over shared memory -> update -> [publish].
//   * slices of one colour go to different warps (boundary slices first), so in a pass every active
//     warp has exactly one slice; the warps with nothing to update in a colour poll the halo instead.
//   * warps whose slice is interior (no halo neighbour) skip the halo barrier altogether.
//   * the reference's "converged?" test (see kernels.cuh, mcgs_kernel) is decided by the end-of-pass
//     barrier itself (barrier.red.or): ONE node whose own residual row exceeds the threshold proves
//     "not converged".  Only a CTA in which no node can prove it takes the slow path of the old kernel.
//   * MAILBOXES instead of a node-indexed exchange array: every part has one slot per halo node in its own
//     halo order (partition.hpp, plan_mailboxes), a boundary node is stored into the slot of every part
//     that reads it (first two slot ids in registers), so the reader's polls are coalesced.  A word is 16 bytes:
//     x, y, z and the pass tag of a node in ONE store and ONE polling load (mb_store / mb_load below; "flag in data"
//     as in mcgs_resident_f32.cuh: no fence, no flag), double-buffered by sweep parity; a slot on another GPU is
//     written through the peer mapping (st.relaxed.sys), never read.
//   * x_ref of the part's nodes is staged in the (not yet needed) matrix-value region of shared memory
//     while r0 = b - A x_ref is formed with the exact fp64 matrix values streamed once; the fp32 values
//     arrive by TMA bulk copy afterwards.
// Pins, obstacles and the outcome of the convergence test are those of mcgs_resident_f32.cuh; the
// shared-memory layout is ResidentPlan::layout(mode 1).
#pragma once
#include "mcgs_resident_f32.cuh"

namespace admmb200 {

#define ADMMB200_OWNED_MAX_COLORS 16

struct OwnedSlice {   // registers; everything but l / rb / ia / dst0 / dst1 is warp-uniform
	int meta;         // -1: none; else colour | boundary << 8 | pinned << 9 | readers << 10 (pinned, readers: per lane)
	int r0, r1;       // ELL rows [r0, r1)
	int l;            // local node id of this lane, -1: padding lane
	unsigned int dst0, dst1; // first two mailbox slots this node is published to (slot | rank << 27); count in meta bits 10..15
	float rb[3];      // r0 = b - A x_ref
	float ia[3];      // 1 / a_ii
};

// One lane's row of the sliced ELL: rows [r0, r1) in batches of 8 with all 24 loads of a batch in flight.
// The last batch is padded by re-reading row r1 - 1 with a zero coefficient instead of a serial tail loop
// (a tail of up to 7 dependent col -> d loads used to cost as much as the three full batches before it).
__device__ __forceinline__ void owned_gather(const float *__restrict__ s_val, const uint16_t *__restrict__ s_col, const float4 *__restrict__ s_d,
	int r0, int r1, int lane, float &sx, float &sy, float &sz)
{
	sx = 0.f; sy = 0.f; sz = 0.f;
	const float *v = s_val + r0 * 32 + lane;
	const uint16_t *c = s_col + r0 * 32 + lane;
	const int n = r1 - r0;
	for (int r = 0; r < n; r += 8) {
		int cc[8]; float a[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int rr = min(r + j, n - 1);
			cc[j] = c[rr * 32];
			a[j] = (r + j < n) ? v[rr * 32] : 0.f;
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const float4 dv = s_d[cc[j]];
			sx = fmaf(a[j], dv.x, sx); sy = fmaf(a[j], dv.y, sy); sz = fmaf(a[j], dv.z, sz);
		}
	}
}

// One mailbox word of this kernel: the three increments of a node and the pass tag in 16 bytes, written and polled with ONE
// 128-bit access.  tools/micro/store_bench.cu (profiles/r02t_store_bench.txt): a publishing warp pays ~25 cycles per store
// instruction plus 4-6 per 128-byte line it touches, and four warps publishing at once share that rate, so three tagged
// 8-byte words in three arrays per node and reader (round 1) cost three times the instructions and lines of one 16-byte
// word; the polls shrink from three loads per slot to one.  Measured: 0.380 -> 0.357 ms per solve on the 1M-tet beam,
// 0.281 -> 0.269 on the 100k-tet beam (same build otherwise).  PTX only promises single-copy atomicity per 64-bit element of the vector, so each half
// carries its own 16-bit tag and is valid on its own: {x, y_hi16 | tag} and {z, y_lo16 | tag}.  tag = solve sequence
// (7 bits) << 9 | pass + 1 (< 512): a slot is rewritten every sweep of every solve, so a stale word can never carry the
// tag a reader is waiting for.
__device__ __forceinline__ void mb_store(ulonglong2 *p, float x, float y, float z, unsigned int tag16, bool peer) {
	const unsigned int yb = __float_as_uint(y);
	const unsigned long long a = (unsigned long long)__float_as_uint(x) | ((unsigned long long)((yb & 0xffff0000u) | tag16) << 32);
	const unsigned long long b = (unsigned long long)__float_as_uint(z) | ((unsigned long long)((yb << 16) | tag16) << 32);
	if (peer) asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
	else asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ bool mb_load(const ulonglong2 *p, unsigned int tag16, float4 &out) {
	unsigned long long a, b;
	asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
	const unsigned int ah = (unsigned int)(a >> 32), bh = (unsigned int)(b >> 32);
	out = make_float4(__uint_as_float((unsigned int)a), __uint_as_float((ah & 0xffff0000u) | (bh >> 16)), __uint_as_float((unsigned int)b), 0.f);
	return (ah & 0xffffu) == tag16 && (bh & 0xffffu) == tag16;
}

__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) :: "memory"); return t; }
__device__ __forceinline__ long long clk_ordered() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) :: "memory"); return t; }

template <int NT, int KMAX, bool OBST, bool PROF>
__global__ void __launch_bounds__(NT, 1) mcgs_owned_f32_kernel(McgsRes32Params R)
{
	constexpr int NW = NT / 32;
	extern __shared__ __align__(128) unsigned char smem[];
	__shared__ double red[32];
	__shared__ __align__(8) uint64_t tma_bar;
	__shared__ short s_role[ADMMB200_OWNED_MAX_COLORS * NW]; // rank of warp w among the pollers of colour c, -1: not a poller
	__shared__ short s_npoll[ADMMB200_OWNED_MAX_COLORS];
	__shared__ int s_decision;
	const McgsParams &P = R.base;
	const long long t_kernel = PROF ? clock64() : 0;
	const PartDesc d = R.parts[R.part0 + blockIdx.x];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n_loc = d.n_own + d.n_halo, C = P.n_colors;

	// shared-memory layout: ResidentPlan::layout(mode 1); the slice tables are not staged (registers instead)
	size_t o = 0;
	auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 15) & ~(size_t)15; return at; };
	float4 *s_d = (float4 *)(smem + take(16 * (size_t)n_loc));
	float *s_val = (float *)(smem + take(res32_val_region((size_t)d.n_rows, (size_t)n_loc)));
	double *s_x = (double *)s_val; // x_ref of owned + halo nodes until r0 is formed; the matrix values arrive afterwards
	uint16_t *s_col = (uint16_t *)(smem + take(sizeof(uint16_t) * 32 * (size_t)d.n_rows));
	int *s_gid = (int *)(smem + take(sizeof(int) * (size_t)n_loc));
	take(sizeof(int) * ((size_t)d.n_slices + 1));
	take(sizeof(short) * (size_t)32 * d.n_slices);
	int *s_cslice = (int *)(smem + take(sizeof(int) * (2 * (size_t)C + 1)));
	int *s_hcol = (int *)(smem + take(sizeof(int) * ((size_t)C + 1)));

	// ---- stage the part: column indices by TMA bulk copy now (the values follow once r0 is formed and
	// their region is free again), index tables by plain loads, d = 0, x_ref into the value region ----
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
		const int g = R.gid[d.gid_off + i];
		s_gid[i] = g; s_d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
		// L2-coherent loads: other parts (and peer GPUs) write P.x at the end of this same kernel and of the previous
		// solve; the read-only path is not allowed to see memory that changes during the kernel
		const double4 xg = ld_node_cg(&P.x[g]);
		s_x[3 * i] = xg.x; s_x[3 * i + 1] = xg.y; s_x[3 * i + 2] = xg.z;
	}
	for (int i = tid; i <= 2 * C; i += NT) s_cslice[i] = R.color_slice[d.cslice_off + i];
	for (int i = tid; i <= C; i += NT) s_hcol[i] = R.halo_color[d.hcolor_off + i];
	__syncthreads();

	// ---- ownership: position pos = k NW + warp in the order "colour by colour, boundary slices first" ----
	auto locate = [&](int pos, int &c, bool &bnd, int &sl) -> bool {
		if (pos >= d.n_slices) return false;
		c = 0;
		while (c + 1 < C && pos >= s_cslice[2 * c + 2]) ++c;
		const int s0 = s_cslice[2 * c], s1 = s_cslice[2 * c + 1], s2 = s_cslice[2 * c + 2];
		const int j = pos - s0, nb = s2 - s1;
		bnd = j < nb;
		sl = bnd ? s1 + j : s0 + (j - nb);
		return true;
	};
	OwnedSlice S[KMAX];
#pragma unroll
	for (int k = 0; k < KMAX; ++k) {
		S[k].meta = -1; S[k].r0 = 0; S[k].r1 = 0; S[k].l = -1; S[k].dst0 = 0u; S[k].dst1 = 0u;
		S[k].rb[0] = S[k].rb[1] = S[k].rb[2] = 0.f; S[k].ia[0] = S[k].ia[1] = S[k].ia[2] = 0.f;
		int c, sl; bool bnd;
		if (locate(k * NW + warp, c, bnd, sl)) {
			S[k].meta = c | (bnd ? 0x100 : 0);
			// a warp's SECOND boundary slice of one colour (parts with more than NW boundary slices in a colour): bit 16
#pragma unroll
			for (int j = 0; j < k; ++j) if (bnd && S[j].meta >= 0 && (S[j].meta & 0x1ff) == (c | 0x100)) S[k].meta |= 0x10000;
			S[k].r0 = __ldg(&R.slice_row[d.slice_off + sl]);
			S[k].r1 = __ldg(&R.slice_row[d.slice_off + sl + 1]);
			S[k].l = (int)__ldg(&R.slice_node[d.snode_off + sl * 32 + lane]);
		}
	}
	// who polls the halo in the pass of colour c: every warp without an interior slice of that colour
	// (those have nothing better to do or need the halo themselves); warp 0 if there is no such warp
	for (int t = tid; t < C * NW; t += NT) {
		const int c = t / NW, w = t % NW;
		bool has_b = false, has_i = false;
		for (int k = 0; k < KMAX; ++k) {
			int cc, sl; bool bnd;
			if (locate(k * NW + w, cc, bnd, sl) && cc == c) { if (bnd) has_b = true; else has_i = true; }
		}
		s_role[t] = (has_b || !has_i) ? 1 : 0;
	}
	if (d.n_rows > 0) mbar_wait(&tma_bar, 0);
	__syncthreads();
	if (tid < C) {
		short rank = 0;
		for (int w = 0; w < NW; ++w) { short &q = s_role[tid * NW + w]; q = q ? rank++ : (short)-1; }
		if (rank == 0) { s_role[tid * NW] = 0; rank = 1; } // somebody has to keep the halo current
		s_npoll[tid] = rank;
	}

	const long long t_staged = PROF ? clock64() : 0;
	unsigned int bar_target = 0;
	const bool check = P.tol2 > 0.0;
	const float omega = (float)P.omega, one_m_omega = (float)(1.0 - P.omega), lb_scale = (float)(1.0 / P.omega - 1.0);
	float4 *pinbuf = R.nodebuf + 2 * (size_t)d.own_off; // [l]: pin - x_ref of a pinned node (read in the first sweep only)
	double xr[OBST ? KMAX : 1][3];                      // x_ref of this lane's nodes: the obstacle test needs absolute positions

	// ---- r0 = b - A x_ref in fp64 with the EXACT matrix (see mcgs_resident_f32.cuh), and |b|^2.  The fp64
	// values stream from global memory once (coalesced); x_ref is gathered from shared memory. ----
	{
		double b2 = 0;
		const double *g_val64 = R.val64 + d.ent_off;
#pragma unroll
		for (int k = 0; k < KMAX; ++k) {
			if (S[k].meta < 0) continue;
			double sx = 0, sy = 0, sz = 0;
			const double *gv = g_val64 + (size_t)S[k].r0 * 32 + lane;
			const uint16_t *cc = s_col + S[k].r0 * 32 + lane;
			const int n = S[k].r1 - S[k].r0;
#pragma unroll 8
			for (int r = 0; r < n; ++r) {
				const double a = __ldg(gv + (size_t)r * 32);
				const double *xc = s_x + 3 * (int)cc[r * 32];
				sx += a * xc[0]; sy += a * xc[1]; sz += a * xc[2];
			}
			const int l = S[k].l;
			if (l >= 0) {
				const int node = s_gid[l];
				const double4 bi = ld_node(&P.b[node]);
				const double xi0 = s_x[3 * l], xi1 = s_x[3 * l + 1], xi2 = s_x[3 * l + 2];
				const double a0 = __ldg(&P.diag[3 * node]), a1 = __ldg(&P.diag[3 * node + 1]), a2 = __ldg(&P.diag[3 * node + 2]);
				const int ps = P.has_pins ? __ldg(&P.pin_slot[node]) : -1;
				if (S[k].meta & 0x100) {
					// where this boundary node is published: one mailbox slot per part that reads it
					const int e0 = __ldg(&R.dest_off[d.own_off + l]), cnt = min(__ldg(&R.dest_off[d.own_off + l + 1]) - e0, 63);
					if (cnt > 0) S[k].dst0 = __ldg(&R.dest_slot[e0]);
					if (cnt > 1) S[k].dst1 = __ldg(&R.dest_slot[e0 + 1]);
					S[k].meta |= cnt << 10;
				}
				S[k].rb[0] = (float)(bi.x - sx - a0 * xi0); S[k].rb[1] = (float)(bi.y - sy - a1 * xi1); S[k].rb[2] = (float)(bi.z - sz - a2 * xi2);
				S[k].ia[0] = (float)(1.0 / a0); S[k].ia[1] = (float)(1.0 / a1); S[k].ia[2] = (float)(1.0 / a2);
				if (OBST) { xr[OBST ? k : 0][0] = xi0; xr[OBST ? k : 0][1] = xi1; xr[OBST ? k : 0][2] = xi2; }
				if (ps >= 0) {
					// pinned node (src/NodalMultiColorGS.hpp:111-117): x = pin, i.e. d = pin - x_ref from its first update on
					S[k].meta |= 0x200;
					pinbuf[l] = make_float4((float)(P.pin_pos[3 * ps] - xi0), (float)(P.pin_pos[3 * ps + 1] - xi1), (float)(P.pin_pos[3 * ps + 2] - xi2), 0.f);
				}
				b2 += bi.x * bi.x + bi.y * bi.y + bi.z * bi.z;
			}
		}
		__syncthreads(); // nobody reads x_ref from the value region any more: the fp32 values may land there
		if (tid == 0 && d.n_rows > 0) {
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			mbar_expect_tx(&tma_bar, val_bytes);
			const unsigned char *gv = (const unsigned char *)(R.val + d.ent_off);
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
	const size_t buf_stride = (size_t)R.total_slots; // mailboxes: [sweep parity][slot] of 16-byte words
	const unsigned int tag_seq = ((R.tag_base >> 12) & 0x7fu) << 9;
	ulonglong2 *const mbox = (ulonglong2 *)R.dglob;
	long long pw = 0, pc = 0, pb = 0, po = 0, t_prev_end = 0, ps1 = 0, ps2 = 0, n_retry = 0, n_spin = 0, hop_nbr = 0, hop_own = 0, hop_first = 0, hop_n = 0;

	// Pulls the halo values of colour `cp` published with tag `tag` in buffer `buf` into shared memory
	// (thread t0 of nt takes every nt-th node).
	auto refresh = [&](int cp, const ulonglong2 *buf, unsigned int tag, int t0, int nt) {
		const int end = s_hcol[cp + 1];
		for (int h = s_hcol[cp] + t0; h < end; h += nt) {
			const ulonglong2 *w = buf + (size_t)d.slot_off + h; // consecutive lanes, consecutive slots: coalesced polls
			float4 v;
			while (!mb_load(w, tag, v)) { if (PROF) ++n_spin; }
			s_d[d.n_own + h] = v;
		}
	};

	// PROF: timeline of passes 40..43 of one part, lane 0 of every warp: trace[(warp * 4 + pass - 40) * 8 + event]
#define TR(ev) do { if (lane == 0 && blockIdx.x == (unsigned)(R.dbg >> 8) && pass >= 40 && pass < 44) R.prof[16 * gridDim.x + ((warp * 4 + (pass - 40)) * 8 + (ev))] = (unsigned long long)clk_ordered(); } while (0)
	int it = 0;
	bool converged = false;
	unsigned int pass = 0; // passes done so far
	const long long t_begin = PROF ? clock64() : 0;
	for (; it < P.iters; ++it) {
		float lb = 0.f; // this lane's part of the lower bound of |b - A x|^2 (rows of the last colour)
		const size_t pub_off = (size_t)(it & 1) * buf_stride;
		for (int color = 0; color < C; ++color) {
			const bool last = check && (color == C - 1);
			const unsigned int pass_tag = tag_seq | (pass + 1);
			const int role = s_role[color * NW + warp];
			long long t0 = 0, t1 = 0;
			if (PROF) { t0 = clk_ordered(); if (t_prev_end) po += t0 - t_prev_end; TR(0); }
			// readers 3 and 4 of a corner node: their slots come from L2 -- asked for now, used after the gather
			// (prefetched for the FIRST boundary slice of this colour the warp owns; a part with more than NW
			// boundary slices of one colour gives a warp a second one, which loads its slots after its gather instead)
			int pre_e0 = 0; unsigned int pre2 = 0u, pre3 = 0u;
#pragma unroll
			for (int k = 0; k < KMAX; ++k) {
				if (S[k].meta < 0 || (S[k].meta & 0x101ff) != (color | 0x100) || S[k].l < 0) continue;
				const int cnt = (S[k].meta >> 10) & 63;
				if (cnt > 2) {
					pre_e0 = __ldg(&R.dest_off[d.own_off + S[k].l]);
					pre2 = __ldg(&R.dest_slot[pre_e0 + 2]);
					if (cnt > 3) pre3 = __ldg(&R.dest_slot[pre_e0 + 3]);
				}
			}
			if (role >= 0) {
				const int n_poll = 32 * (int)s_npoll[color];
				if (pass > 0 && !(PROF && (R.dbg & 2))) {
					// what the neighbours changed in the previous pass: the halo nodes of that pass's colour
					const int cp = (color + C - 1) % C;
					const int it_prev = color > 0 ? it : it - 1;
					refresh(cp, mbox + (size_t)(it_prev & 1) * buf_stride, tag_seq | pass, 32 * role + lane, n_poll);
				}
				if (PROF) TR(1);
				named_sync(1, n_poll);
				if (PROF && (R.dbg & 128) && pass > 0 && role == 0 && lane == 0) {
					// when did the neighbours publish what has just arrived?  (global timer; single GPU only)
					const unsigned long long now = gtime_ns();
					volatile unsigned long long *pubt = (volatile unsigned long long *)(R.prof + 16 * gridDim.x + 1024);
					unsigned long long latest = 0, earliest = ~0ull;
					for (int i = 0; i < d.n_nbr; ++i) {
						const unsigned long long v = pubt[(size_t)__ldg(&R.nbr[d.nbr_off + i]) * 128 + ((pass - 1) & 127)];
						if (v != 0 && v <= now && now - v < 100000ull) { latest = v > latest ? v : latest; earliest = v < earliest ? v : earliest; }
					}
					const unsigned long long own = pubt[(size_t)(R.part0 + blockIdx.x) * 128 + ((pass - 1) & 127)];
					if (latest != 0 && own != 0 && own <= now) { hop_nbr += (long long)(now - latest); hop_first += (long long)(now - earliest); hop_own += (long long)(now - own); ++hop_n; }
				}
			}
			if (PROF) { t1 = clk_ordered(); TR(2); }
#pragma unroll
			for (int k = 0; k < KMAX; ++k) {
				if ((S[k].meta & 0xff) != color || S[k].meta < 0) continue;
				if (PROF && (R.dbg & 4) && !(S[k].meta & 0x100)) continue; // timing experiment: no interior work
				float sx, sy, sz;
				owned_gather(s_val, s_col, s_d, S[k].r0, (PROF && (R.dbg & 32)) ? S[k].r0 : S[k].r1, lane, sx, sy, sz);
				if (PROF) TR(5);
				const int l = S[k].l;
				if (l < 0) continue;
				const float4 dold = s_d[l];
				float4 dn;
				if (S[k].meta & 0x200) {
					dn = dold;
					if (it == 0) dn = pinbuf[l];
				} else {
					// segment_update (src/NodalMultiColorGS.hpp:180-215) on the increment
					const float g0 = (S[k].rb[0] - sx) * S[k].ia[0], g1 = (S[k].rb[1] - sy) * S[k].ia[1], g2 = (S[k].rb[2] - sz) * S[k].ia[2];
					dn = make_float4(fmaf(omega, g0, one_m_omega * dold.x), fmaf(omega, g1, one_m_omega * dold.y), fmaf(omega, g2, one_m_omega * dold.z), 0.f);
					bool hit = false;
					if (OBST) {
						const double *x0 = xr[OBST ? k : 0];
						double gs[3] = {x0[0] + (double)g0, x0[1] + (double)g1, x0[2] + (double)g2};
						double nx[3] = {x0[0] + (double)dn.x, x0[1] + (double)dn.y, x0[2] + (double)dn.z};
						hit = mcgs_collide(P.obs, P.n_obstacles, gs, nx);
						if (hit) dn = make_float4((float)(nx[0] - x0[0]), (float)(nx[1] - x0[1]), (float)(nx[2] - x0[2]), 0.f);
					}
					if (last && !hit) {
						// residual row right after the update: a_ii (1/omega - 1) (d_new - d_old); steering only (4x margin)
						const float rx = lb_scale * __fdividef(dn.x - dold.x, S[k].ia[0]), ry = lb_scale * __fdividef(dn.y - dold.y, S[k].ia[1]), rz = lb_scale * __fdividef(dn.z - dold.z, S[k].ia[2]);
						lb += rx * rx + ry * ry + rz * rz;
					}
				}
				s_d[l] = dn;
				if (PROF) TR(6);
				if ((S[k].meta & 0x100) && !(PROF && (R.dbg & 16))) {
					const int cnt = (S[k].meta >> 10) & 63;
					auto put = [&](unsigned int ent) {
						const unsigned int q = ent >> 27;
						const size_t at = pub_off + (size_t)(ent & 0x7ffffffu);
						if ((int)q == R.rank) mb_store(mbox + at, dn.x, dn.y, dn.z, pass_tag, false);
						else mb_store((ulonglong2 *)R.peer_dglob[q] + at, dn.x, dn.y, dn.z, pass_tag, true);
					};
					if (cnt > 0) put(S[k].dst0);
					if (cnt > 1) put(S[k].dst1);
					if (cnt > 2) { // a corner node read by more than two parts
						int e0 = pre_e0; unsigned int p2 = pre2, p3 = pre3;
						if (S[k].meta & 0x10000) { e0 = __ldg(&R.dest_off[d.own_off + l]); p2 = __ldg(&R.dest_slot[e0 + 2]); if (cnt > 3) p3 = __ldg(&R.dest_slot[e0 + 3]); }
						put(p2);
						if (cnt > 3) put(p3);
						for (int e = e0 + 4; e < e0 + cnt; ++e) put(__ldg(&R.dest_slot[e]));
					}
					if (PROF && (R.dbg & 64)) __threadfence(); // timing experiment: does a fence get the published words out sooner?
					if (PROF && (R.dbg & 128) && lane == 0 && R.world == 1) atomicMax(R.prof + 16 * gridDim.x + 1024 + (size_t)(R.part0 + blockIdx.x) * 128 + (pass & 127), gtime_ns());
				}
			}
			long long t2 = 0;
			if (PROF) t2 = clk_ordered();
			if (PROF) { pc += t2 - t1; pw += t1 - t0; TR(3); }
			++pass;
			if (!last) { __syncthreads(); if (PROF) { t_prev_end = clk_ordered(); pb += t_prev_end - t2; --pass; TR(4); ++pass; } continue; }

			// ---- "converged?" after the sweep (see kernels.cuh).  The barrier that ends the pass also answers
			// "can any lane of this part prove |b - A x|^2 >= 4 tol^2 |b|^2 from its own rows alone?" ----
			const int proven = __syncthreads_or((double)lb >= thresh && lb > 0.f);
			if (PROF) { t_prev_end = clk_ordered(); pb += t_prev_end - t2; --pass; TR(4); ++pass; }
			if (proven) {
				// leave a note for parts that cannot prove it themselves; nobody waits here.  ONE atomic carries both
				// "arrived" (low 16 bits) and "proved it" (high bits): a waiting part reads a single word, so it can never
				// see all arrivals without the proof that came with one of them (a separate flag store could be overtaken)
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
				refresh(C - 1, mbox + pub_off, tag_seq | pass, tid, NT);
				__syncthreads();
				double acc = 0;
#pragma unroll
				for (int k = 0; k < KMAX; ++k) {
					if (S[k].meta < 0) continue;
					float sx, sy, sz;
					owned_gather(s_val, s_col, s_d, S[k].r0, S[k].r1, lane, sx, sy, sz);
					const int l = S[k].l;
					if (l >= 0) {
						const float4 dv = s_d[l];
						double rx = (double)S[k].rb[0] - (double)sx - (double)dv.x / (double)S[k].ia[0];
						double ry = (double)S[k].rb[1] - (double)sy - (double)dv.y / (double)S[k].ia[1];
						double rz = (double)S[k].rb[2] - (double)sz - (double)dv.z / (double)S[k].ia[2];
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
		const int node = s_gid[l];
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
		q[0] = (unsigned long long)pw; q[1] = (unsigned long long)pc; q[2] = (unsigned long long)pb; q[3] = (unsigned long long)(clock64() - t_kernel); q[5] = (unsigned long long)po; q[6] = (unsigned long long)ps1; q[7] = (unsigned long long)ps2; q[8] = (unsigned long long)n_spin; q[9] = (unsigned long long)n_retry;
		q[13] = (unsigned long long)(t_staged - t_kernel); q[14] = (unsigned long long)(t_begin - t_staged); q[15] = (unsigned long long)(t_loop_end - t_begin);
	}
	if (PROF && lane == 0 && hop_n) { unsigned long long *q = R.prof + 16 * blockIdx.x; q[10] = (unsigned long long)hop_nbr; q[11] = (unsigned long long)hop_own; q[12] = (unsigned long long)hop_n; q[4] = (unsigned long long)hop_first; }
	(void)hop_nbr; (void)hop_own; (void)hop_first; (void)hop_n;
	(void)pw; (void)pc; (void)pb; (void)po; (void)t_prev_end; (void)ps1; (void)ps2; (void)n_retry; (void)n_spin;
}

} // namespace admmb200
