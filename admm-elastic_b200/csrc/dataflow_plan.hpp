// dataflow_plan.hpp -- host-side planning (and a host-side model) of a BARRIER-FREE schedule for the
// shared-memory-resident multi-colour Gauss-Seidel.  NEXT STEP of DESIGN.md 9.1: planned and checked on the
// host here; no kernel consumes it yet.
//
// Today (mcgs_owned_f32.cuh) every colour pass of a part ends with a CTA barrier and starts with "all halo
// values of the previous colour have arrived", so a part moves at the pace of its slowest slice and of the
// slowest of its ~13 neighbours, 120 times per solve.  The sweep itself only needs, for a slice S (32 nodes
// of one colour) in sweep `it`:
//   * every OWN slice T that holds a neighbour of a node of S:  done[T] >= it + (colour(T) < colour(S))
//   * every HALO node h that a row of S reads: the value its owner published in pass
//       (it, colour(h)) if colour(h) < colour(S), else (it - 1, colour(h))          (nothing for it = 0)
// and, because the matrix pattern is symmetric, the same two conditions also guarantee that nobody still
// reads the values S is about to overwrite (the readers of S's old values are exactly the slices / parts S
// reads from, one pass later in the cycle).  With per-slice counters in shared memory and tagged mailboxes
// (already there) no barrier is needed: interior slices run ahead as far as their neighbours allow, a
// boundary slice waits for the two or three parts it really touches.
//
// plan_dataflow derives the dependency lists from a ResidentPlan; simulate_dataflow executes the plan on the
// host with a randomised scheduler that honours ONLY these conditions (plus the program order of the warp
// that owns a slice) and compares with plain colour-by-colour sweeps: bit-identical or it throws.
#pragma once
#include "partition.hpp"
#include <random>

namespace admmb200 {

struct DataflowPlan {
	int n_warps = 16;
	// per part: slices in the order "colour by colour, boundary first" (as mcgs_owned_f32_kernel assigns them:
	// position pos -> warp pos % n_warps, turn pos / n_warps)
	struct Part {
		std::vector<int> order;                 // pos -> slice id of the ResidentPlan part
		std::vector<int> color;                 // per slice id
		std::vector<std::vector<int>> deps;     // per slice id: own slices it reads from (= that read from it)
		std::vector<std::vector<int>> halo;     // per slice id: halo indices h (0 .. n_halo) its rows read
	};
	std::vector<Part> parts;
	size_t max_deps = 0, max_halo_refs = 0;
};

inline DataflowPlan plan_dataflow(const ResidentPlan &R, int n_colors, int n_warps = 16)
{
	if (R.lanes != 1) throw std::runtime_error("dataflow plan: one lane per node only");
	DataflowPlan D;
	D.n_warps = n_warps;
	D.parts.resize(R.parts.size());
	for (size_t p = 0; p < R.parts.size(); ++p) {
		const PartDesc &d = R.parts[p];
		DataflowPlan::Part &P = D.parts[p];
		const int *cs = R.color_slice.data() + d.cslice_off;
		const int *srow = R.slice_row.data() + d.slice_off;
		const short *snode = R.slice_node.data() + d.snode_off;
		P.color.assign(d.n_slices, 0);
		for (int c = 0; c < n_colors; ++c) {
			for (int sl = cs[2 * c]; sl < cs[2 * c + 2]; ++sl) P.color[sl] = c;
			for (int sl = cs[2 * c + 1]; sl < cs[2 * c + 2]; ++sl) P.order.push_back(sl); // boundary slices first
			for (int sl = cs[2 * c]; sl < cs[2 * c + 1]; ++sl) P.order.push_back(sl);
		}
		std::vector<int> slice_of(d.n_own, -1);
		for (int sl = 0; sl < d.n_slices; ++sl) for (int g = 0; g < 32; ++g) { const int l = snode[sl * 32 + g]; if (l >= 0) slice_of[l] = sl; }
		P.deps.assign(d.n_slices, {});
		P.halo.assign(d.n_slices, {});
		for (int sl = 0; sl < d.n_slices; ++sl) {
			std::vector<int> &dep = P.deps[sl], &hl = P.halo[sl];
			for (int r = srow[sl]; r < srow[sl + 1]; ++r) for (int g = 0; g < 32; ++g) {
				const size_t e = (size_t)d.ent_off + (size_t)r * 32 + g;
				if (R.val[e] == 0.0) continue; // padding
				const int cl = R.col[e];
				if (cl < d.n_own) { if (slice_of[cl] != sl) dep.push_back(slice_of[cl]); }
				else hl.push_back(cl - d.n_own);
			}
			std::sort(dep.begin(), dep.end()); dep.erase(std::unique(dep.begin(), dep.end()), dep.end());
			std::sort(hl.begin(), hl.end()); hl.erase(std::unique(hl.begin(), hl.end()), hl.end());
			for (int t : dep) if (P.color[t] == P.color[sl]) throw std::runtime_error("dataflow plan: two slices of one colour depend on each other");
			D.max_deps = std::max(D.max_deps, dep.size());
			D.max_halo_refs = std::max(D.max_halo_refs, hl.size());
		}
	}
	return D;
}

// Host model.  diag[node], rhs[node]: one scalar system (the three components are independent); omega: SOR.
// Returns the number of scheduling decisions taken.  Throws when the dataflow execution deadlocks or differs
// from the colour-by-colour reference in any bit.
inline long long simulate_dataflow(const ResidentPlan &R, const DataflowPlan &D, int n_nodes, int n_colors, const double *diag, const double *rhs,
	double omega, int sweeps, unsigned int seed, double *x_out)
{
	const int n_parts = (int)R.parts.size();
	// ---- reference: sweeps x colours over all parts, rows summed in the plan's order ----
	std::vector<double> xr((size_t)n_nodes, 0.0);
	auto row_update = [&](const PartDesc &d, int sl, int g, const std::vector<double> &own_and_halo) -> double {
		const int *srow = R.slice_row.data() + d.slice_off;
		const short *snode = R.slice_node.data() + d.snode_off;
		const int l = snode[sl * 32 + g];
		double s = 0;
		for (int r = srow[sl]; r < srow[sl + 1]; ++r) { const size_t e = (size_t)d.ent_off + (size_t)r * 32 + g; s += R.val[e] * own_and_halo[R.col[e]]; }
		const int node = R.gid[d.gid_off + l];
		return (1.0 - omega) * own_and_halo[l] + omega * (rhs[node] - s) / diag[node];
	};
	{
		std::vector<double> loc;
		for (int it = 0; it < sweeps; ++it) for (int c = 0; c < n_colors; ++c) {
			std::vector<std::pair<int, double>> upd;
			for (int p = 0; p < n_parts; ++p) {
				const PartDesc &d = R.parts[p];
				const int *cs = R.color_slice.data() + d.cslice_off;
				const short *snode = R.slice_node.data() + d.snode_off;
				loc.resize((size_t)d.n_own + d.n_halo);
				for (int i = 0; i < d.n_own + d.n_halo; ++i) loc[i] = xr[R.gid[d.gid_off + i]];
				for (int sl = cs[2 * c]; sl < cs[2 * c + 2]; ++sl) for (int g = 0; g < 32; ++g) {
					const int l = snode[sl * 32 + g];
					if (l >= 0) upd.emplace_back(R.gid[d.gid_off + l], row_update(d, sl, g, loc));
				}
			}
			for (auto &u : upd) xr[u.first] = u.second;
		}
	}
	// ---- dataflow execution ----
	struct Slot { double v[2]; int pass[2]; };           // mailbox word per (reader part, halo index): double-buffered by sweep parity
	std::vector<std::vector<Slot>> mbox(n_parts);
	std::vector<std::vector<double>> loc(n_parts);      // shared-memory image: own + halo values
	std::vector<std::vector<int>> done(n_parts);         // sweeps completed per slice
	std::vector<std::vector<int>> turn(n_parts);         // per warp: index of its next task in its program
	std::vector<std::vector<std::vector<std::pair<int, int>>>> prog(n_parts); // per warp: (sweep, slice) in program order
	std::vector<int> color_of((size_t)n_nodes, -1);
	for (int p = 0; p < n_parts; ++p) {
		const PartDesc &d = R.parts[p];
		mbox[p].assign(d.n_halo, Slot{{0, 0}, {-1, -1}});
		loc[p].assign((size_t)d.n_own + d.n_halo, 0.0);
		done[p].assign(d.n_slices, 0);
		turn[p].assign(D.n_warps, 0);
		prog[p].assign(D.n_warps, {});
		const DataflowPlan::Part &P = D.parts[p];
		for (int it = 0; it < sweeps; ++it) for (size_t pos = 0; pos < P.order.size(); ++pos) prog[p][pos % D.n_warps].emplace_back(it, P.order[pos]);
		// a warp runs its slices of one sweep in colour order
		for (auto &w : prog[p]) std::stable_sort(w.begin(), w.end(), [&](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first != b.first ? a.first < b.first : P.color[a.second] < P.color[b.second]; });
		const short *snode = R.slice_node.data() + d.snode_off;
		for (int sl = 0; sl < d.n_slices; ++sl) for (int g = 0; g < 32; ++g) { const int l = snode[sl * 32 + g]; if (l >= 0) color_of[R.gid[d.gid_off + l]] = P.color[sl]; }
	}
	// readers of a node: (part, halo index)
	std::vector<std::vector<std::pair<int, int>>> readers((size_t)n_nodes);
	for (int p = 0; p < n_parts; ++p) { const PartDesc &d = R.parts[p]; for (int h = 0; h < d.n_halo; ++h) readers[R.gid[d.gid_off + d.n_own + h]].emplace_back(p, h); }

	std::mt19937 rng(seed);
	long long decisions = 0, remaining = 0;
	for (int p = 0; p < n_parts; ++p) for (auto &w : prog[p]) remaining += (long long)w.size();
	std::vector<std::pair<int, int>> ready;
	while (remaining > 0) {
		ready.clear();
		for (int p = 0; p < n_parts; ++p) {
			const PartDesc &d = R.parts[p];
			const DataflowPlan::Part &P = D.parts[p];
			for (int w = 0; w < D.n_warps; ++w) {
				if (turn[p][w] >= (int)prog[p][w].size()) continue;
				const int it = prog[p][w][turn[p][w]].first, sl = prog[p][w][turn[p][w]].second, c = P.color[sl];
				bool ok = done[p][sl] == it;
				for (int t : P.deps[sl]) if (ok && done[p][t] < it + (P.color[t] < c ? 1 : 0)) ok = false;
				for (int h : P.halo[sl]) {
					if (!ok) break;
					const int ch = color_of[R.gid[d.gid_off + d.n_own + h]];
					const int its = ch < c ? it : it - 1;          // sweep in which the value was produced
					if (its < 0) continue;                          // nothing published yet: the initial zero is right
					if (mbox[p][h].pass[its & 1] != its * n_colors + ch) ok = false;
				}
				if (ok) ready.emplace_back(p, w);
			}
		}
		if (ready.empty()) throw std::runtime_error("dataflow model: deadlock");
		const std::pair<int, int> pick = ready[rng() % ready.size()];
		const int p = pick.first, w = pick.second;
		const PartDesc &d = R.parts[p];
		const DataflowPlan::Part &P = D.parts[p];
		const int it = prog[p][w][turn[p][w]].first, sl = prog[p][w][turn[p][w]].second, c = P.color[sl];
		// pull the halo values this slice reads (each with the tag its colour implies), gather, update, publish
		for (int h : P.halo[sl]) {
			const int ch = color_of[R.gid[d.gid_off + d.n_own + h]];
			const int its = ch < c ? it : it - 1;
			if (its >= 0) loc[p][d.n_own + h] = mbox[p][h].v[its & 1];
		}
		const short *snode = R.slice_node.data() + d.snode_off;
		double nv[32];
		for (int g = 0; g < 32; ++g) if (snode[sl * 32 + g] >= 0) nv[g] = row_update(d, sl, g, loc[p]);
		for (int g = 0; g < 32; ++g) {
			const int l = snode[sl * 32 + g];
			if (l < 0) continue;
			loc[p][l] = nv[g];
			for (auto &rd : readers[R.gid[d.gid_off + l]]) { Slot &s = mbox[rd.first][rd.second]; s.v[it & 1] = nv[g]; s.pass[it & 1] = it * n_colors + c; }
		}
		done[p][sl] = it + 1;
		++turn[p][w];
		--remaining; ++decisions;
	}
	for (int p = 0; p < n_parts; ++p) { const PartDesc &d = R.parts[p]; for (int l = 0; l < d.n_own; ++l) { const int node = R.gid[d.gid_off + l]; if (loc[p][l] != xr[node]) throw std::runtime_error("dataflow model: result differs from the colour-by-colour sweeps"); if (x_out) x_out[node] = loc[p][l]; } }
	return decisions;
}

} // namespace admmb200
