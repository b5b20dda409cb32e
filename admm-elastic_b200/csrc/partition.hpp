// partition.hpp -- host-side planning of the shared-memory-resident multi-colour Gauss-Seidel.
//
// The constant system matrix of one ADMM solve sequence (A = L (x) I3 + M, src/Solver.cpp:226) is cut
// into one part per SM.  A part owns a spatially compact set of nodes (recursive coordinate bisection
// on the rest positions, balanced by row length), keeps the rows of those nodes AND their current
// positions in the SM's shared memory for all sweeps x colours of a solve, and only touches L2/HBM for
// the halo (neighbours owned by other parts), the right-hand side and the result.
// Colours stay global (graphcolor::color_matrix semantics, deps/mclscene/include/MCL/GraphColor.hpp:
// 66-72): all parts sweep the same colour between two grid barriers, so the update order -- and with
// it the result -- is the same as in the reference's NodalMultiColorGS::solve
// (src/NodalMultiColorGS.hpp:96-131) for the same colouring.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace admmb200 {

#ifdef __CUDACC__
#define ADMMB200_HD __host__ __device__
#else
#define ADMMB200_HD
#endif
// fp32 resident solve: the matrix-value region doubles as the place where the anchor positions x_ref of
// the part's nodes (3 doubles, owned + halo) sit while r0 = b - A x_ref is formed, before the values arrive
ADMMB200_HD inline size_t res32_val_region(size_t n_rows, size_t n_loc) { const size_t a = 4 * 32 * n_rows, b = 24 * n_loc; return a > b ? a : b; }

struct PartDesc {
	int n_own, n_halo, n_slices, n_rows; // owned nodes, halo nodes, slices (warps' work items), 32-entry ELL rows
	long long ent_off;                   // first entry of this part in the global col/val arrays
	int gid_off;                         // into gid[]: n_own owned ids (local order) then n_halo halo ids
	int slice_off;                       // into slice_row[] (n_slices + 1 entries, relative row numbers)
	int snode_off;                       // into slice_node[] (n_slices * G local ids, -1 = padding)
	int cslice_off;                      // into color_slice[] (2 n_colors + 1 entries: per colour first interior
	                                     // slice, first boundary slice; then the end)
	int nbr_off, n_nbr;                  // into nbr[]: parts this part exchanges halo values with
	int own_off;                         // number of nodes owned by the parts before this one
	int hcolor_off;                      // into halo_color[] (n_colors + 1 entries): halo nodes are sorted by colour
	int slot_off;                        // first mailbox slot of this part: slot_off + h receives the value of halo node h
	int pad_[1];
};

struct ResidentPlan {
	int n_parts = 0, lanes = 4;
	std::vector<PartDesc> parts;
	std::vector<uint16_t> col;     // local index: < n_own -> shared-memory x, else halo (gid[col])
	std::vector<double> val;       // converted to the storage precision at upload
	std::vector<int> gid, slice_row, color_slice, nbr, halo_color;
	std::vector<short> slice_node;
	std::vector<int> part_of;      // node -> part (for tests / diagnostics)
	// Mailboxes (mcgs_owned_f32.cuh): every part has one slot per halo node, in its own halo order, so the
	// reader's polls are coalesced.  A boundary node is written into the slot of every part that reads it:
	// dest_off[own_off + l] .. dest_off[own_off + l + 1] indexes dest_slot[], entry = slot | rank << 27.
	std::vector<int> dest_off;
	std::vector<unsigned int> dest_slot;
	size_t total_slots = 0;
	size_t max_nbr = 0;
	bool schedule_banks = true;   // T = 1: order the entries of each row for conflict-free float4 gathers (detail::schedule_slice)
	long long cycles_before = 0, cycles_after = 0; // modelled shared-memory cycles of all gathers of one sweep, before / after that ordering
	size_t max_rows = 0, max_own = 0, max_halo = 0, max_slices = 0, entries = 0, nnz = 0;
	// dynamic shared memory one CTA needs with `val_bytes`-wide matrix values
	// dynamic shared memory one CTA needs.  mode 0: positions of the owned nodes as 3 doubles, matrix
	// values `val_bytes` wide (mcgs_resident_kernel); mode 1: float4 increments of owned AND halo nodes,
	// float values (mcgs_resident_f32_kernel, mcgs_owned_f32_kernel); mode 2: the same for mcgs_tiled_f32_kernel.
	size_t smem_bytes(int n_colors, int val_bytes, int mode = 0) const {
		size_t worst = 0;
		for (const PartDesc &d : parts) worst = std::max(worst, layout(d, n_colors, val_bytes, lanes, mode, nullptr));
		return worst;
	}
	// byte offsets of the shared-memory arrays of one part (the kernels compute the same)
	static size_t layout(const PartDesc &d, int n_colors, int val_bytes, int lanes, int mode, size_t *off /* [8] or null */) {
		const int G = 32 / lanes; // nodes per slice
		size_t o = 0, tmp[8];
		auto take = [&](int i, size_t bytes) { tmp[i] = o; o += (bytes + 15) & ~(size_t)15; };
		if (mode == 0) take(0, sizeof(double) * 3 * (size_t)d.n_own);               // x of owned nodes
		else take(0, 16 * ((size_t)d.n_own + d.n_halo));                             // float4 increments, owned + halo
		take(1, mode == 0 ? (size_t)val_bytes * 32 * (size_t)d.n_rows : res32_val_region((size_t)d.n_rows, (size_t)d.n_own + d.n_halo)); // val
		take(2, sizeof(uint16_t) * 32 * (size_t)d.n_rows);                           // col
		if (mode == 2) {
			// mcgs_tiled_f32.cuh: no node ids and no per-lane slice tables in shared memory, but the scaled right-hand side
			take(3, sizeof(float) * 3 * (size_t)d.n_own);                            // rbs = omega r0 / a_ii
			take(4, sizeof(int) * ((size_t)d.n_slices + 1));                         // slice_row
			take(5, sizeof(int) * (size_t)d.n_slices);                               // first node | valid nodes << 16
		} else {
		take(3, sizeof(int) * ((size_t)d.n_own + d.n_halo));                         // gid
		take(4, sizeof(int) * ((size_t)d.n_slices + 1));                             // slice_row
		take(5, sizeof(short) * (size_t)G * d.n_slices);                             // slice_node
		}
		take(6, sizeof(int) * (2 * (size_t)n_colors + 1));                           // color_slice
		take(7, mode == 0 ? 0 : sizeof(int) * ((size_t)n_colors + 1));               // halo_color
		if (off) std::memcpy(off, tmp, sizeof(tmp));
		return o;
	}
};

namespace detail {
// ---------------------------------------------------------------------------------------------
// Bank-conflict-aware ordering of the entries of a slice (one lane per node, T = 1).
//
// The sweep gathers `float4 d[col]` from shared memory: one LDS.128 per ELL row-step.  A 128-bit warp load is served in
// four quarter-warp phases (8 lanes x 16 B = the 128-byte bank width); a phase takes as many cycles as the most loaded
// 16-byte bank group (col mod 8) has DISTINCT words.  With the entries of a row in matrix order the 8 columns of a phase
// are effectively random: ~2.8 cycles per phase, 11-13 per row-step (tools/micro/gather_bench.cu: 285 cycles per
// 18-row slice at saturation where 4 per row-step would give ~110).  The order of the entries inside a row is free (a
// row sum), and padding entries (zero coefficient) may point at any node, so each quarter-warp's entries are scheduled
// here like an edge colouring of the bipartite multigraph lanes x bank groups: per row-step a matching that gives every
// lane an entry of a bank group nobody else in its quarter uses in that step.  Lanes with the least slack (remaining
// entries == remaining steps) are served first, augmenting paths move earlier choices out of the way, bank groups with
// many remaining entries are preferred.  What cannot be matched stays a conflict (counted by slice_conflict_cycles).
// ---------------------------------------------------------------------------------------------
struct SliceEntry { uint16_t col; double val; };

// cycles of the four phases of one row-step: per quarter the largest number of distinct words in one bank group
inline int rowstep_cycles(const uint16_t *c32)
{
	int total = 0;
	for (int q = 0; q < 4; ++q) {
		int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
		for (int a = 0; a < 8; ++a) {
			bool dup = false;
			for (int b = 0; b < a; ++b) if (c32[8 * q + b] == c32[8 * q + a]) dup = true;
			if (!dup) ++cnt[c32[8 * q + a] & 7];
		}
		int m = 1;
		for (int g = 0; g < 8; ++g) m = std::max(m, cnt[g]);
		total += m;
	}
	return total;
}

// rows[lane] = the real entries of that lane's row (any order); width = row-steps of the slice; self[lane] = an owned
// local index a padding entry may fall back to; n_loc = local nodes (padding may point at any of them).
// Writes col/val [width][32].
inline void schedule_slice(std::vector<SliceEntry> rows[32], int width, const int *self, int n_loc, uint16_t *col, double *val)
{
	for (int q = 0; q < 4; ++q) {
		std::vector<SliceEntry> *R = rows + 8 * q;
		int used[8]; // entries of lane a already placed
		std::vector<char> taken[8];
		for (int a = 0; a < 8; ++a) { used[a] = 0; taken[a].assign(R[a].size(), 0); }
		for (int j = 0; j < width; ++j) {
			const int rs = width - j;
			int gdeg[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // remaining entries per bank group (over the quarter)
			for (int a = 0; a < 8; ++a) for (size_t e = 0; e < R[a].size(); ++e) if (!taken[a][e]) ++gdeg[R[a][e].col & 7];
			int lane_of_group[8], pick[8]; // matching: group -> lane, lane -> entry index (-1: none yet)
			for (int g = 0; g < 8; ++g) lane_of_group[g] = -1;
			for (int a = 0; a < 8; ++a) pick[a] = -1;
			int order[8];
			for (int a = 0; a < 8; ++a) order[a] = a;
			std::sort(order, order + 8, [&](int a, int b) {
				const int sa = rs - ((int)R[a].size() - used[a]), sb = rs - ((int)R[b].size() - used[b]);
				return sa != sb ? sa < sb : a < b; // least slack first
			});
			// augmenting-path search: lane a looks for a free group among its remaining entries
			std::function<bool(int, char *)> augment = [&](int a, char *seen) -> bool {
				// candidate groups of lane a, most loaded first
				int cand[8], nc = 0;
				for (size_t e = 0; e < R[a].size(); ++e) if (!taken[a][e]) {
					const int g = R[a][e].col & 7;
					bool have = false;
					for (int i = 0; i < nc; ++i) if (cand[i] == g) have = true;
					if (!have) cand[nc++] = g;
				}
				std::sort(cand, cand + nc, [&](int x, int y) { return gdeg[x] != gdeg[y] ? gdeg[x] > gdeg[y] : x < y; });
				for (int i = 0; i < nc; ++i) {
					const int g = cand[i];
					if (seen[g]) continue;
					seen[g] = 1;
					if (lane_of_group[g] < 0 || augment(lane_of_group[g], seen)) {
						lane_of_group[g] = a;
						for (size_t e = 0; e < R[a].size(); ++e) if (!taken[a][e] && (R[a][e].col & 7) == g) { pick[a] = (int)e; break; }
						return true;
					}
				}
				return false;
			};
			for (int oi = 0; oi < 8; ++oi) {
				const int a = order[oi];
				if ((int)R[a].size() - used[a] <= 0) continue;
				char seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
				augment(a, seen);
			}
			// lanes without a match: pad if they can, otherwise place a conflicting entry (least loaded group this step)
			int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
			for (int a = 0; a < 8; ++a) if (pick[a] >= 0) ++load[R[a][pick[a]].col & 7];
			for (int oi = 0; oi < 8; ++oi) {
				const int a = order[oi];
				const int left = (int)R[a].size() - used[a];
				if (pick[a] >= 0 || left <= 0 || left < rs) continue; // matched, or nothing left, or may pad
				int best = -1;
				for (size_t e = 0; e < R[a].size(); ++e) if (!taken[a][e]) { if (best < 0 || load[R[a][e].col & 7] < load[R[a][best].col & 7]) best = (int)e; }
				pick[a] = best;
				++load[R[a][best].col & 7];
			}
			for (int a = 0; a < 8; ++a) {
				const size_t at = (size_t)j * 32 + 8 * q + a;
				if (pick[a] >= 0) {
					col[at] = R[a][pick[a]].col; val[at] = R[a][pick[a]].val;
					taken[a][pick[a]] = 1; ++used[a];
				} else {
					// padding: zero coefficient, any word of a bank group nobody uses in this step
					int g = self[8 * q + a] & 7;
					if (load[g] > 0) for (int t = 0; t < 8; ++t) if (load[t] == 0) { g = t; break; }
					int c = (self[8 * q + a] & ~7) | g;
					if (c >= n_loc) c = g < n_loc ? g : self[8 * q + a];
					col[at] = (uint16_t)c; val[at] = 0.0;
					++load[g];
				}
			}
		}
		for (int a = 0; a < 8; ++a) if (used[a] != (int)R[a].size()) throw std::runtime_error("resident plan: slice scheduling lost an entry");
	}
}

inline void rcb(std::vector<int> &ids, int lo, int hi, int k, int part0, const double *pos, const std::vector<double> &w, std::vector<int> &part_of)
{
	if (k <= 1 || hi - lo <= 0) { for (int i = lo; i < hi; ++i) part_of[ids[i]] = part0; return; }
	double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300}, total = 0;
	for (int i = lo; i < hi; ++i) {
		const double *p = pos + 3 * (size_t)ids[i];
		for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], p[a]); mx[a] = std::max(mx[a], p[a]); }
		total += w[ids[i]];
	}
	int axis = 0;
	for (int a = 1; a < 3; ++a) if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
	std::sort(ids.begin() + lo, ids.begin() + hi, [&](int a, int b) {
		double pa = pos[3 * (size_t)a + axis], pb = pos[3 * (size_t)b + axis];
		return pa < pb || (pa == pb && a < b);
	});
	const int kl = k / 2, kr = k - kl;
	const double target = total * (double)kl / (double)k;
	double acc = 0;
	int mid = lo;
	while (mid < hi && acc + 0.5 * w[ids[mid]] < target) { acc += w[ids[mid]]; ++mid; }
	// keep at least one node per side when there are enough nodes
	if (hi - lo >= 2) mid = std::min(std::max(mid, lo + 1), hi - 1);
	rcb(ids, lo, mid, kl, part0, pos, w, part_of);
	rcb(ids, mid, hi, kr, part0 + kl, pos, w, part_of);
}
} // namespace detail

// node -> part: recursive coordinate bisection balanced by padded row length (+ the node's own update)
inline std::vector<int> plan_partition(int n, const int *rowptr, const int *cols, const double *vals, const double *pos3, int n_parts, int lanes)
{
	// The weights do not depend on the lane count of the sweep kernel: the host (multi-GPU element selection,
	// admm_b200_plan_parts) and finalize must arrive at the SAME partition whichever kernel finalize picks.
	(void)lanes;
	std::vector<double> w(n);
	for (int i = 0; i < n; ++i) {
		int len = 0;
		for (int q = rowptr[i]; q < rowptr[i + 1]; ++q) if (cols[q] != i && vals[q] != 0.0) ++len;
		w[i] = (double)((len + 3) / 4 * 4) + 2.0;
	}
	std::vector<int> part_of(n, 0), ids(n);
	std::iota(ids.begin(), ids.end(), 0);
	detail::rcb(ids, 0, n, n_parts, 0, pos3, w, part_of);
	return part_of;
}

// rowptr/cols/vals: scalar matrix L (both triangles); color_of_list: colour -> node lists (offsets, nodes)
inline ResidentPlan plan_resident(int n, const int *rowptr, const int *cols, const double *vals, int n_colors, const int *color_off,
	const int *color_nodes, const double *pos3, int n_parts, int lanes = 4)
{
	ResidentPlan R;
	R.n_parts = n_parts; R.lanes = lanes;
	const int T = lanes, G = 32 / T;
	std::vector<int> color_of(n, -1);
	for (int c = 0; c < n_colors; ++c) for (int k = color_off[c]; k < color_off[c + 1]; ++k) color_of[color_nodes[k]] = c;
	std::vector<int> rowlen(n, 0);
	for (int i = 0; i < n; ++i) {
		int len = 0;
		for (int q = rowptr[i]; q < rowptr[i + 1]; ++q) if (cols[q] != i && vals[q] != 0.0) ++len;
		rowlen[i] = len;
	}
	R.part_of = plan_partition(n, rowptr, cols, vals, pos3, n_parts, lanes);

	std::vector<std::vector<int>> own(n_parts);
	for (int i = 0; i < n; ++i) own[R.part_of[i]].push_back(i);
	// boundary node = has a neighbour owned by another part (the pattern is symmetric, so it is also
	// exactly the set of nodes other parts read as halo)
	std::vector<char> boundary(n, 0);
	for (int i = 0; i < n; ++i)
		for (int q = rowptr[i]; q < rowptr[i + 1]; ++q)
			if (cols[q] != i && vals[q] != 0.0 && R.part_of[cols[q]] != R.part_of[i]) { boundary[i] = 1; boundary[cols[q]] = 1; }
	std::vector<int> local_of(n, -1);
	R.parts.resize(n_parts);
	int own_running = 0;
	// colour-major, long rows first inside a colour (uniform slices -> little ELL padding), ids last
	// and inside a colour the interior nodes first: they are updated while the neighbours' flags of
	// the previous pass are still in flight
	std::vector<int> owner_local(n, 0); // position of a node in its owner's order
	for (int p = 0; p < n_parts; ++p) {
		std::vector<int> &nodes = own[p];
		std::sort(nodes.begin(), nodes.end(), [&](int a, int b) {
			if (color_of[a] != color_of[b]) return color_of[a] < color_of[b];
			if (boundary[a] != boundary[b]) return boundary[a] < boundary[b];
			if (rowlen[a] != rowlen[b]) return rowlen[a] > rowlen[b];
			return a < b;
		});
		for (size_t l = 0; l < nodes.size(); ++l) owner_local[nodes[l]] = (int)l;
	}
	for (int p = 0; p < n_parts; ++p) {
		std::vector<int> &nodes = own[p];
		PartDesc d;
		std::memset(&d, 0, sizeof(d));
		d.n_own = (int)nodes.size();
		d.own_off = own_running;
		own_running += d.n_own;
		d.gid_off = (int)R.gid.size();
		d.slice_off = (int)R.slice_row.size();
		d.snode_off = (int)R.slice_node.size();
		d.cslice_off = (int)R.color_slice.size();
		d.ent_off = (long long)R.col.size();
		for (int l = 0; l < d.n_own; ++l) { local_of[nodes[l]] = l; R.gid.push_back(nodes[l]); }
		// halo nodes, sorted by colour: after a pass only the halo nodes of that pass's colour have new
		// values, and they are contiguous ...
		std::vector<int> halo;
		for (int l = 0; l < d.n_own; ++l) {
			const int node = nodes[l];
			for (int q = rowptr[node]; q < rowptr[node + 1]; ++q) {
				const int g = cols[q];
				if (g == node || vals[q] == 0.0 || R.part_of[g] == p || local_of[g] == -2) continue;
				local_of[g] = -2;
				halo.push_back(g);
			}
		}
		// ... and inside a colour by owner part and the owner's own order: the nodes one neighbour publishes
		// to this part then sit in consecutive mailbox slots, so its stores coalesce (plan_mailboxes)
		std::sort(halo.begin(), halo.end(), [&](int a, int b) {
			if (color_of[a] != color_of[b]) return color_of[a] < color_of[b];
			if (R.part_of[a] != R.part_of[b]) return R.part_of[a] < R.part_of[b];
			return owner_local[a] < owner_local[b];
		});
		d.hcolor_off = (int)R.halo_color.size();
		{
			size_t h = 0;
			for (int c = 0; c < n_colors; ++c) { R.halo_color.push_back((int)h); while (h < halo.size() && color_of[halo[h]] == c) ++h; }
			R.halo_color.push_back((int)halo.size());
		}
		for (size_t h = 0; h < halo.size(); ++h) local_of[halo[h]] = d.n_own + (int)h;
		auto local_index = [&](int g) -> int { return local_of[g]; };
		int rows = 0, slices = 0;
		R.slice_row.push_back(0);
		int k = 0;
		for (int cc = 0; cc < 2 * n_colors; ++cc) {
			const int c = cc / 2, bnd = cc % 2;
			R.color_slice.push_back(slices);
			int k1 = k;
			while (k1 < d.n_own && color_of[nodes[k1]] == c && boundary[nodes[k1]] == bnd) ++k1;
			for (; k < k1; k += G) {
				int width = 0;
				for (int g = 0; g < G && k + g < k1; ++g) width = std::max(width, (rowlen[nodes[k + g]] + T - 1) / T);
				size_t base = R.col.size();
				R.col.resize(base + (size_t)width * 32, 0);
				R.val.resize(base + (size_t)width * 32, 0.0);
				for (int g = 0; g < G; ++g) {
					int node = (k + g < k1) ? nodes[k + g] : -1;
					R.slice_node.push_back(node < 0 ? (short)-1 : (short)(k + g));
					int self = node < 0 ? (k < k1 ? k : 0) : k + g; // padding entries read an owned node with a zero coefficient
					int j = 0;
					if (node >= 0) {
						for (int q = rowptr[node]; q < rowptr[node + 1]; ++q) {
							int cg = cols[q];
							if (cg == node || vals[q] == 0.0) continue;
							int li = local_index(cg);
							if (li > 65535) throw std::runtime_error("resident plan: part too large for 16-bit local indices");
							R.col[base + (size_t)(j / T) * 32 + g * T + (j % T)] = (uint16_t)li;
							R.val[base + (size_t)(j / T) * 32 + g * T + (j % T)] = vals[q];
							++j; ++R.nnz;
						}
					}
					for (; j < width * T; ++j) R.col[base + (size_t)(j / T) * 32 + g * T + (j % T)] = (uint16_t)self;
				}
				if (T == 1 && R.schedule_banks && width > 0) {
					// re-order the entries of every lane's row so that the quarter-warp phases of the float4 gather hit distinct
					// shared-memory bank groups (detail::schedule_slice)
					std::vector<detail::SliceEntry> lane_rows[32];
					int self[32];
					for (int g = 0; g < 32; ++g) {
						const int node = (k + g < k1) ? nodes[k + g] : -1;
						self[g] = node < 0 ? (k < k1 ? k : 0) : k + g;
						if (node < 0) continue;
						for (int j = 0; j < rowlen[node]; ++j) lane_rows[g].push_back({R.col[base + (size_t)j * 32 + g], R.val[base + (size_t)j * 32 + g]});
					}
					R.cycles_before += [&]() { long long c = 0; for (int j = 0; j < width; ++j) c += detail::rowstep_cycles(&R.col[base + (size_t)j * 32]); return c; }();
					detail::schedule_slice(lane_rows, width, self, d.n_own + (int)halo.size(), &R.col[base], &R.val[base]);
					R.cycles_after += [&]() { long long c = 0; for (int j = 0; j < width; ++j) c += detail::rowstep_cycles(&R.col[base + (size_t)j * 32]); return c; }();
				}
				rows += width;
				++slices;
				R.slice_row.push_back(rows);
			}
			k = k1;
		}
		R.color_slice.push_back(slices);
		d.n_halo = (int)halo.size();
		d.n_slices = slices;
		d.n_rows = rows;
		for (int g : halo) { R.gid.push_back(g); local_of[g] = -1; }
		for (int l = 0; l < d.n_own; ++l) local_of[nodes[l]] = -1;
		if (d.n_own > 32767) throw std::runtime_error("resident plan: part too large for 16-bit slice nodes");
		R.parts[p] = d;
		R.max_rows = std::max(R.max_rows, (size_t)rows);
		R.max_own = std::max(R.max_own, (size_t)d.n_own);
		R.max_halo = std::max(R.max_halo, (size_t)d.n_halo);
		R.max_slices = std::max(R.max_slices, (size_t)slices);
	}
	// neighbour parts (symmetric): q is a neighbour of p when either reads halo values of the other
	std::vector<std::vector<int>> nb(n_parts);
	for (int p = 0; p < n_parts; ++p) {
		const PartDesc &d = R.parts[p];
		for (int h = 0; h < d.n_halo; ++h) {
			int q = R.part_of[R.gid[d.gid_off + d.n_own + h]];
			nb[p].push_back(q);
			nb[q].push_back(p);
		}
	}
	for (int p = 0; p < n_parts; ++p) {
		std::sort(nb[p].begin(), nb[p].end());
		nb[p].erase(std::unique(nb[p].begin(), nb[p].end()), nb[p].end());
		R.parts[p].nbr_off = (int)R.nbr.size();
		R.parts[p].n_nbr = (int)nb[p].size();
		R.nbr.insert(R.nbr.end(), nb[p].begin(), nb[p].end());
		R.max_nbr = std::max(R.max_nbr, nb[p].size());
	}
	R.entries = R.col.size();
	return R;
}

// Fills the mailbox tables of a plan.  parts_per_rank: parts [r * parts_per_rank, ...) run on rank r (n_parts for one GPU).
inline void plan_mailboxes(ResidentPlan &R, int n_nodes, int parts_per_rank)
{
	size_t slots = 0;
	for (PartDesc &d : R.parts) { d.slot_off = (int)slots; slots += (size_t)d.n_halo; }
	if (slots >= ((size_t)1 << 27)) throw std::runtime_error("resident plan: too many halo slots");
	R.total_slots = slots;
	// node -> position in the owner-ordered numbering (own_off + local id)
	std::vector<int> owner_pos((size_t)n_nodes, -1);
	for (const PartDesc &d : R.parts) for (int l = 0; l < d.n_own; ++l) owner_pos[R.gid[d.gid_off + l]] = d.own_off + l;
	std::vector<int> count((size_t)n_nodes + 1, 0);
	for (const PartDesc &d : R.parts) for (int h = 0; h < d.n_halo; ++h) ++count[owner_pos[R.gid[d.gid_off + d.n_own + h]] + 1];
	R.dest_off.assign((size_t)n_nodes + 1, 0);
	for (int i = 0; i < n_nodes; ++i) R.dest_off[i + 1] = R.dest_off[i] + count[i + 1];
	R.dest_slot.assign(slots ? slots : 1, 0u);
	std::vector<int> fill(R.dest_off.begin(), R.dest_off.end() - 1);
	for (int p = 0; p < R.n_parts; ++p) {
		const PartDesc &d = R.parts[p];
		const unsigned int rank = (unsigned int)(p / parts_per_rank);
		for (int h = 0; h < d.n_halo; ++h) {
			const int pos = owner_pos[R.gid[d.gid_off + d.n_own + h]];
			R.dest_slot[fill[pos]++] = (unsigned int)(d.slot_off + h) | (rank << 27);
		}
	}
}

// Multi-GPU: parts [r * parts_per_rank, (r+1) * parts_per_rank) live on rank r.  For every node owned by
// `rank`, the set of OTHER ranks that read it as halo (bit q = rank q).  The pattern is symmetric, so this
// is also the set of ranks whose values this rank reads next to that node.
inline std::vector<unsigned int> dest_masks(const ResidentPlan &R, int n_nodes, int parts_per_rank, int rank)
{
	std::vector<unsigned int> mask((size_t)n_nodes, 0u);
	for (int p = 0; p < R.n_parts; ++p) {
		const int q = p / parts_per_rank;
		if (q == rank) continue;
		const PartDesc &d = R.parts[p];
		for (int h = 0; h < d.n_halo; ++h) {
			const int g = R.gid[d.gid_off + d.n_own + h];
			if (R.part_of[g] / parts_per_rank == rank) mask[g] |= 1u << q;
		}
	}
	return mask;
}

} // namespace admmb200
