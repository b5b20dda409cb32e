// sparse.hpp -- host-side (init-time) sparse helpers of the B200 ADMM-elastic solver:
//   * assembly of the scalar system matrix L with A = M + dt^2 D^T W^2 D = L (x) I3 + M
//     (reference: src/Solver.cpp:207-226 builds the 3n x 3n matrix with Eigen products)
//   * node colouring for NodalMultiColorGS (reference: mcl::graphcolor::color_matrix,
//     deps/mclscene/include/MCL/GraphColor.hpp:66-253 -- randomised and time-seeded there, so any
//     valid colouring is an equally legitimate input to the sweep)
//   * fill-reducing ordering + L D L^T factorisation handed to the GPU triangular solves
//     (reference: Eigen::SimplicialLDLT::compute in LDLTSolver::update_system,
//     src/LinearSolver.hpp:79-84)
// None of this is on the per-iteration hot path; it runs once in Solver::initialize.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace admm_b200 {
namespace sparse {

struct Csr {
	int n = 0;
	std::vector<int> rowptr, cols;
	std::vector<double> vals;
};

struct Entry { int row, col; double val; };

// Sums duplicate (row, col) entries in a fixed order (sorted by row, col, then insertion order) so
// the result is reproducible run to run.
inline Csr from_entries(int n, std::vector<Entry> &e)
{
	std::stable_sort(e.begin(), e.end(), [](const Entry &a, const Entry &b) { return a.row != b.row ? a.row < b.row : a.col < b.col; });
	Csr m; m.n = n; m.rowptr.assign(n + 1, 0);
	for (size_t i = 0; i < e.size();) {
		size_t j = i; double s = 0;
		while (j < e.size() && e[j].row == e[i].row && e[j].col == e[i].col) { s += e[j].val; ++j; }
		m.cols.push_back(e[i].col); m.vals.push_back(s); m.rowptr[e[i].row + 1]++;
		i = j;
	}
	for (int i = 0; i < n; ++i) m.rowptr[i + 1] += m.rowptr[i];
	return m;
}

// Greedy colouring in node order: smallest colour not used by an already coloured neighbour.
// Output: colour -> ascending node list, like graphcolor::make_map.
inline void color_greedy(const Csr &A, std::vector<std::vector<int>> &colors)
{
	// Largest-degree-first greedy (Welsh-Powell), ties by node id: deterministic, and on tet meshes it
	// gives few, evenly filled colours (4 on the block beams where index order gives 9, five of them
	// nearly empty) -- every colour costs one synchronisation per sweep on the GPU.
	const int n = A.n;
	std::vector<int> deg(n, 0), order(n);
	for (int i = 0; i < n; ++i) for (int q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q) if (A.cols[q] != i && A.vals[q] != 0.0) ++deg[i];
	std::iota(order.begin(), order.end(), 0);
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return deg[a] > deg[b]; });
	std::vector<int> color(n, -1), mark;
	int n_colors = 0;
	for (int i : order) {
		mark.assign(n_colors + 1, 0);
		for (int q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q) {
			int j = A.cols[q];
			if (j != i && A.vals[q] != 0.0 && color[j] >= 0) mark[color[j]] = 1;
		}
		int c = 0;
		while (mark[c]) ++c;
		color[i] = c;
		n_colors = std::max(n_colors, c + 1);
	}
	colors.assign(n_colors, std::vector<int>());
	for (int i = 0; i < n; ++i) colors[color[i]].push_back(i);
}

// Randomised palette colouring in the spirit of Grable & Panconesi as used by the reference
// (GraphColor.hpp:119-231): every uncoloured node draws a colour from its palette, keeps it when no
// neighbour drew the same, palettes shrink by the colours neighbours fixed and grow when they run
// dry.  Deterministic (own LCG, fixed seed); produces about as many colours as the reference does.
inline void color_random_palette(const Csr &A, std::vector<std::vector<int>> &colors, uint32_t seed = 1)
{
	const int n = A.n;
	const int init_palette = 6;
	std::vector<std::vector<int>> palette(n);
	for (int i = 0; i < n; ++i) { palette[i].resize(init_palette); std::iota(palette[i].begin(), palette[i].end(), 0); }
	std::vector<int> color(n, -1), trial(n, -1), queue(n);
	std::iota(queue.begin(), queue.end(), 0);
	uint64_t state = seed * 6364136223846793005ULL + 1442695040888963407ULL;
	auto rnd = [&]() { state = state * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(state >> 33); };
	for (int round = 0; !queue.empty(); ++round) {
		if (round > n + 64) throw std::runtime_error("graphcolor::color Error: Nodes remain uncolored");
		for (int i : queue) trial[i] = palette[i][rnd() % palette[i].size()];
		std::vector<int> next;
		std::vector<char> keep(queue.size(), 1);
		for (size_t k = 0; k < queue.size(); ++k) {
			int i = queue[k];
			for (int q = A.rowptr[i]; q < A.rowptr[i + 1] && keep[k]; ++q) {
				int j = A.cols[q];
				if (j == i || A.vals[q] == 0.0) continue;
				if (color[j] == trial[i]) keep[k] = 0;                  // neighbour already owns it
				else if (color[j] < 0 && trial[j] == trial[i] && j > i) keep[k] = 0; // tie: larger index keeps
			}
		}
		for (size_t k = 0; k < queue.size(); ++k) if (keep[k]) color[queue[k]] = trial[queue[k]];
		for (size_t k = 0; k < queue.size(); ++k) {
			int i = queue[k];
			if (keep[k]) continue;
			std::vector<int> &pal = palette[i];
			for (int q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q) {
				int j = A.cols[q];
				if (j == i || A.vals[q] == 0.0 || color[j] < 0) continue;
				pal.erase(std::remove(pal.begin(), pal.end(), color[j]), pal.end());
			}
			if (pal.size() < 2) pal.push_back(init_palette + round);
			next.push_back(i);
		}
		queue.swap(next);
	}
	int n_colors = 0;
	for (int i = 0; i < n; ++i) n_colors = std::max(n_colors, color[i] + 1);
	std::vector<std::vector<int>> all(n_colors);
	for (int i = 0; i < n; ++i) all[color[i]].push_back(i);
	colors.clear();
	for (auto &c : all) if (!c.empty()) colors.push_back(c);
}

inline bool coloring_is_valid(const Csr &A, const std::vector<std::vector<int>> &colors)
{
	std::vector<int> color(A.n, -1);
	for (size_t c = 0; c < colors.size(); ++c) for (int i : colors[c]) { if (i < 0 || i >= A.n || color[i] >= 0) return false; color[i] = (int)c; }
	for (int i = 0; i < A.n; ++i) {
		if (color[i] < 0) return false;
		for (int q = A.rowptr[i]; q < A.rowptr[i + 1]; ++q) { int j = A.cols[q]; if (j != i && A.vals[q] != 0.0 && color[j] == color[i]) return false; }
	}
	return true;
}

// ---------------------------------------------------------------------------------------------
// ordering: geometric nested dissection (positions are known at initialize time)
// ---------------------------------------------------------------------------------------------
struct NdWork {
	const Csr *A;
	const double *pos; // 3 per node
	std::vector<int> part; // scratch: which side a node is on during a split
	std::vector<int> order;
	int leaf;
};

inline void nd_recurse(NdWork &w, std::vector<int> &nodes)
{
	if ((int)nodes.size() <= w.leaf) { for (int i : nodes) w.order.push_back(i); return; }
	double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
	for (int i : nodes) for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], w.pos[3 * i + d]); hi[d] = std::max(hi[d], w.pos[3 * i + d]); }
	int ax = 0;
	for (int d = 1; d < 3; ++d) if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
	size_t mid = nodes.size() / 2;
	std::nth_element(nodes.begin(), nodes.begin() + mid, nodes.end(), [&](int a, int b) {
		double pa = w.pos[3 * a + ax], pb = w.pos[3 * b + ax];
		return pa != pb ? pa < pb : a < b;
	});
	for (size_t k = 0; k < nodes.size(); ++k) w.part[nodes[k]] = k < mid ? 1 : 2;
	std::vector<int> left, right, sep;
	for (size_t k = 0; k < nodes.size(); ++k) {
		int i = nodes[k];
		if (k < mid) { left.push_back(i); continue; }
		bool touches = false;
		for (int q = w.A->rowptr[i]; q < w.A->rowptr[i + 1] && !touches; ++q) { int j = w.A->cols[q]; if (j != i && w.part[j] == 1) touches = true; }
		(touches ? sep : right).push_back(i);
	}
	for (int i : nodes) w.part[i] = 0;
	nodes.clear(); nodes.shrink_to_fit();
	if (left.empty() || (right.empty() && sep.size() >= mid)) {
		// no useful split (e.g. a clique): stop recursing
		for (int i : left) w.order.push_back(i);
		for (int i : right) w.order.push_back(i);
		for (int i : sep) w.order.push_back(i);
		return;
	}
	nd_recurse(w, left);
	nd_recurse(w, right);
	for (int i : sep) w.order.push_back(i);
}

// perm[new] = old
inline std::vector<int> order_nested_dissection(const Csr &A, const double *pos3, int leaf = 32)
{
	NdWork w; w.A = &A; w.pos = pos3; w.part.assign(A.n, 0); w.leaf = leaf;
	std::vector<int> all(A.n);
	std::iota(all.begin(), all.end(), 0);
	nd_recurse(w, all);
	if ((int)w.order.size() != A.n) throw std::runtime_error("nested dissection lost nodes");
	return w.order;
}

// ---------------------------------------------------------------------------------------------
// L D L^T (up-looking, elimination-tree based) of P A P^T.  A symmetric, both triangles stored.
// Output: unit lower L in CSC without the diagonal, D.
// ---------------------------------------------------------------------------------------------
struct Ldlt {
	int n = 0;
	std::vector<int> perm; // perm[new] = old
	std::vector<int> Lp, Li;
	std::vector<double> Lx, D;
};

inline Ldlt factor_ldlt(const Csr &A, const std::vector<int> &perm)
{
	const int n = A.n;
	Ldlt f; f.n = n; f.perm = perm;
	std::vector<int> pinv(n);
	for (int k = 0; k < n; ++k) pinv[perm[k]] = k;
	// B = P A P^T, upper triangle, column-wise (= lower triangle row-wise of the symmetric matrix)
	std::vector<int> Bp(n + 1, 0), Bi; std::vector<double> Bx;
	for (int k = 0; k < n; ++k) {
		int old = perm[k];
		for (int q = A.rowptr[old]; q < A.rowptr[old + 1]; ++q) if (pinv[A.cols[q]] <= k) Bp[k + 1]++;
	}
	for (int k = 0; k < n; ++k) Bp[k + 1] += Bp[k];
	Bi.resize(Bp[n]); Bx.resize(Bp[n]);
	for (int k = 0; k < n; ++k) {
		int old = perm[k], p = Bp[k];
		for (int q = A.rowptr[old]; q < A.rowptr[old + 1]; ++q) { int i = pinv[A.cols[q]]; if (i <= k) { Bi[p] = i; Bx[p] = A.vals[q]; ++p; } }
	}
	std::vector<int> parent(n, -1), flag(n), lnz(n, 0), pattern(n);
	for (int k = 0; k < n; ++k) {
		flag[k] = k;
		for (int p = Bp[k]; p < Bp[k + 1]; ++p) {
			int i = Bi[p];
			if (i < k) for (; flag[i] != k; i = parent[i]) { if (parent[i] == -1) parent[i] = k; lnz[i]++; flag[i] = k; }
		}
	}
	f.Lp.assign(n + 1, 0);
	for (int k = 0; k < n; ++k) f.Lp[k + 1] = f.Lp[k] + lnz[k];
	f.Li.resize(f.Lp[n]); f.Lx.resize(f.Lp[n]); f.D.assign(n, 0.0);
	std::vector<double> Y(n, 0.0);
	std::fill(lnz.begin(), lnz.end(), 0);
	for (int k = 0; k < n; ++k) {
		int top = n;
		flag[k] = k;
		for (int p = Bp[k]; p < Bp[k + 1]; ++p) {
			int i = Bi[p];
			Y[i] += Bx[p];
			int len = 0;
			for (; flag[i] != k; i = parent[i]) { pattern[len++] = i; flag[i] = k; }
			while (len > 0) pattern[--top] = pattern[--len];
		}
		f.D[k] = Y[k]; Y[k] = 0.0;
		for (; top < n; ++top) {
			int i = pattern[top];
			double yi = Y[i]; Y[i] = 0.0;
			int p2 = f.Lp[i] + lnz[i];
			for (int p = f.Lp[i]; p < p2; ++p) Y[f.Li[p]] -= f.Lx[p] * yi;
			double lki = yi / f.D[i];
			f.D[k] -= lki * yi;
			f.Li[p2] = k; f.Lx[p2] = lki; lnz[i]++;
		}
		if (f.D[k] == 0.0) throw std::runtime_error("**LDLTSolver Error: zero pivot in factorisation");
	}
	return f;
}

} // namespace sparse
} // namespace admm_b200
