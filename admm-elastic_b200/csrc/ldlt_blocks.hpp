// ldlt_blocks.hpp -- host-side plan of the block (supernodal) triangular solves of sptrsv_blocks.cuh.
//
// A level-scheduled sparse triangular solve needs one grid-wide barrier per dependency level, and a nested-dissection
// factor has as many levels as the separators along the deepest path have columns: 1 199 for the 100k-tet beam, 2 415
// for the 512 x 512 cloth -- every separator is a dense triangle whose rows depend on each other one by one (measured:
// 176 / 576 ms per solve, slower than the reference's CPU solve).  Here the columns of L are cut into contiguous BLOCKS:
//   * chains of the elimination tree: column j joins the block of j-1 when parent(j-1) = j, so the diagonal block L_BB is an
//     (almost) dense unit lower triangle -- the separators of the dissection;
//   * consecutive small blocks are merged up to 32 columns (leaves of the dissection).
// Any contiguous cut is valid: with L = [L_BB on the diagonal, L_{B,<B} to the left] forward substitution block by block is
//     y_B = L_BB^-1 ( b_B - L_{B,<B} y_{<B} )
// and L_BB^-1 is precomputed here (dense, unit lower), so inside a block nothing depends on anything: the sparse part is
// a row gather from blocks that are already done, the dense part a matrix-vector product.  Blocks that do not read each
// other form a level; the number of levels is the height of the block tree (about 2 log2(n / 32) instead of thousands).
// The backward solve x_B = L_BB^-T ( D^-1 y_B - L_{>B,B}^T x_{>B} ) uses the same blocks with the transposed inverse.
// Works for ANY factor handed to admm_b200_set_ldlt (Eigen's AMD ordering included); nested dissection keeps it shallow.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <vector>

namespace admmb200 {

struct LdltBlockPlan {
	int n = 0, n_blocks = 0, n_levels_f = 0, n_levels_b = 0;
	std::vector<int> blk_of;          // [n] block of a column
	std::vector<int> blk_c0;          // [n_blocks + 1] first column of a block
	std::vector<long long> inv_off;   // [n_blocks] offset of the block's packed inverse (s (s-1) / 2 entries)
	std::vector<double> inv;          // L_BB^-1, strictly lower part, row r holds columns 0..r-1 at r (r-1) / 2
	std::vector<double> invT;         // its transpose, row r holds columns r+1..s-1 at r (s-1) - r (r-1) / 2
	// forward: rows ordered by the level of their block; entries of a row that lie LEFT of its block (CSR)
	std::vector<int> f_lev_ptr, f_rows, f_rowptr, f_cols;
	std::vector<double> f_vals;
	std::vector<int> f_lanes;         // [2 * n_levels_f] lanes per row in the gather and in the dense phase
	// backward: columns ordered by the backward level of their block; entries of a column BELOW its block (CSC)
	std::vector<int> b_lev_ptr, b_cols, b_colptr, b_rows;
	std::vector<double> b_vals;
	std::vector<int> b_lanes;
	long long nnz_out = 0, nnz_inv = 0;
	int max_block = 0;
	// The bottom forest: blocks below level `cut` form independent subtrees of the block tree; every CTA of the solve
	// kernel owns some of them and walks them level by level with CTA barriers only.  A segment = the rows of one level of
	// one subtree, f_rows[begin, end), executed in this order forward and in reverse order backward; only levels >= cut
	// need grid barriers.  lanes[4 * level + {0, 1, 2, 3}] = threads per row in the forward gather / forward dense /
	// backward gather / backward dense phase.
	int cut = 0, n_ctas = 0;
	std::vector<int> seg_ptr;         // [n_ctas + 1] into the segment arrays
	std::vector<int> seg_begin, seg_end, seg_level;
	std::vector<int> lanes;           // [4 * n_levels_f]
	// Everything a phase needs to know about its k-th row in ONE 16-byte load (the phases are chains of dependent L2 / DRAM
	// round trips; rows -> rowptr -> entries costs two more than this): by position k in f_rows
	std::vector<int> desc_fg;         // [4 n] {row i, first, one past the last entry in f_cols / f_vals, perm[i]}
	std::vector<int> desc_bg;         // [4 n] {column j, first, one past the last entry in b_rows / b_vals, perm[j]}
	std::vector<int> desc_d;          // [4 n] {first column c0 of the block, r = i - c0, block size s, row i}
	std::vector<long long> desc_off;  // [n] offset of the block's packed inverse
	void build_descriptors(const int *perm) {
		desc_fg.resize(4 * (size_t)n); desc_bg.resize(4 * (size_t)n); desc_d.resize(4 * (size_t)n); desc_off.resize((size_t)n);
		for (int k = 0; k < n; ++k) {
			const int i = f_rows[k], b = blk_of[i], c0 = blk_c0[b];
			int *g = &desc_fg[4 * (size_t)k], *h = &desc_bg[4 * (size_t)k], *d = &desc_d[4 * (size_t)k];
			g[0] = i; g[1] = f_rowptr[i]; g[2] = f_rowptr[i + 1]; g[3] = perm[i];
			h[0] = i; h[1] = b_colptr[i]; h[2] = b_colptr[i + 1]; h[3] = perm[i];
			d[0] = c0; d[1] = i - c0; d[2] = blk_c0[b + 1] - c0; d[3] = i;
			desc_off[k] = inv_off[b];
		}
	}
};

// Threads per row of a phase: about 8 entries per thread (4 loads in flight, twice), but never fewer rows in flight than the
// grid has room for -- the rows of the top separators hold thousands of entries while only a few hundred rows exist, so a
// row gets up to a whole CTA (1024 threads, reduced through shared memory in sptrsv_blocks.cuh).
inline int ldlt_pow2_lanes(double avg_len, int n_rows = 1 << 30, int grid_threads = 148 * 1024)
{
	int t = 1;
	while (t < 1024 && (double)t * 8.0 < avg_len) t <<= 1;
	while (t > 1 && (long long)n_rows * t > 2LL * grid_threads) t >>= 1; // keep at most two rounds of rows
	return t;
}

// Lp/Li/Lx: strictly lower unit L in CSC with ascending row indices per column (what sparse::factor_ldlt and Eigen's
// SimplicialLDLT produce).
inline LdltBlockPlan plan_ldlt_blocks(int n, const int *Lp, const int *Li, const double *Lx, int n_ctas = 148, int cta_threads = 1024, int merge_up_to = 32, int max_block = 4096)
{
	LdltBlockPlan P;
	P.n = n;
	// ---- blocks ----
	std::vector<int> start; // first column of every block
	for (int j = 0; j < n; ++j) {
		bool join = false;
		if (j > 0) {
			const int cnt_prev = Lp[j] - Lp[j - 1], cnt = Lp[j + 1] - Lp[j];
			const int parent_prev = cnt_prev > 0 ? Li[Lp[j - 1]] : -1;
			// chain of the elimination tree.  struct(L_{j-1}) \ {j} is a subset of struct(L_j) then; where it is a proper subset
			// (the leading columns of a separator) the diagonal block gets explicit zeros, which the dense inverse absorbs
			(void)cnt;
			join = parent_prev == j && (j - start.back()) < max_block;
		}
		if (!join) start.push_back(j);
	}
	start.push_back(n);
	// merge consecutive small blocks (dissection leaves): the merged diagonal block is treated as dense
	{
		std::vector<int> merged;
		size_t b = 0;
		const size_t nb = start.size() - 1;
		while (b < nb) {
			size_t e = b + 1;
			// only along the elimination tree: block e joins when it holds the parent of the last column merged so far (it
			// depends on that block anyway); merging unrelated neighbours would chain independent subtrees together
			while (e < nb && start[e + 1] - start[b] <= merge_up_to) {
				const int last = start[e] - 1;
				const int par = Lp[last + 1] > Lp[last] ? Li[Lp[last]] : -1;
				if (par < start[e] || par >= start[e + 1]) break;
				++e;
			}
			merged.push_back(start[b]);
			b = e;
		}
		merged.push_back(n);
		start.swap(merged);
	}
	P.n_blocks = (int)start.size() - 1;
	P.blk_c0 = start;
	P.blk_of.assign(n, 0);
	for (int b = 0; b < P.n_blocks; ++b) { for (int j = start[b]; j < start[b + 1]; ++j) P.blk_of[j] = b; P.max_block = std::max(P.max_block, start[b + 1] - start[b]); }

	// ---- dense inverses of the diagonal blocks ----
	P.inv_off.assign(P.n_blocks, 0);
	long long total = 0;
	for (int b = 0; b < P.n_blocks; ++b) { const long long s = start[b + 1] - start[b]; P.inv_off[b] = total; total += s * (s - 1) / 2; }
	P.nnz_inv = total;
	P.inv.assign((size_t)std::max<long long>(total, 1), 0.0);
	P.invT.assign((size_t)std::max<long long>(total, 1), 0.0);
	{
		std::vector<double> M, X;
		for (int b = 0; b < P.n_blocks; ++b) {
			const int c0 = start[b], s = start[b + 1] - c0;
			if (s <= 1) continue;
			M.assign((size_t)s * s, 0.0); // M[i * s + k] = L(c0 + i, c0 + k), i > k
			for (int k = 0; k < s; ++k)
				for (int q = Lp[c0 + k]; q < Lp[c0 + k + 1]; ++q) { const int i = Li[q] - c0; if (i >= s) break; M[(size_t)i * s + k] = Lx[q]; }
			// X = M^-1 (unit lower) by forward substitution on the rows: X_i = e_i - sum_{k<i} M_ik X_k
			X.assign((size_t)s * s, 0.0);
			for (int i = 0; i < s; ++i) {
				double *xi = &X[(size_t)i * s];
				xi[i] = 1.0;
				for (int k = 0; k < i; ++k) {
					const double m = M[(size_t)i * s + k];
					if (m == 0.0) continue;
					const double *xk = &X[(size_t)k * s];
					for (int c = 0; c <= k; ++c) xi[c] -= m * xk[c];
				}
			}
			double *inv = &P.inv[(size_t)P.inv_off[b]], *invT = &P.invT[(size_t)P.inv_off[b]];
			for (int r = 0; r < s; ++r) {
				for (int k = 0; k < r; ++k) inv[(size_t)r * (r - 1) / 2 + k] = X[(size_t)r * s + k];
				const size_t at = (size_t)r * (s - 1) - (size_t)r * (r - 1) / 2;
				for (int k = r + 1; k < s; ++k) invT[at + (k - r - 1)] = X[(size_t)k * s + r];
			}
		}
	}

	// ---- entries outside the diagonal blocks: CSR by row (forward), CSC by column (backward) ----
	P.f_rowptr.assign(n + 1, 0);
	P.b_colptr.assign(n + 1, 0);
	for (int j = 0; j < n; ++j)
		for (int q = Lp[j]; q < Lp[j + 1]; ++q) if (P.blk_of[Li[q]] != P.blk_of[j]) { P.f_rowptr[Li[q] + 1]++; P.b_colptr[j + 1]++; }
	for (int i = 0; i < n; ++i) { P.f_rowptr[i + 1] += P.f_rowptr[i]; P.b_colptr[i + 1] += P.b_colptr[i]; }
	P.nnz_out = P.f_rowptr[n];
	P.f_cols.resize((size_t)P.nnz_out); P.f_vals.resize((size_t)P.nnz_out); P.b_rows.resize((size_t)P.nnz_out); P.b_vals.resize((size_t)P.nnz_out);
	{
		std::vector<int> ffill(P.f_rowptr.begin(), P.f_rowptr.end() - 1), bfill(P.b_colptr.begin(), P.b_colptr.end() - 1);
		for (int j = 0; j < n; ++j)
			for (int q = Lp[j]; q < Lp[j + 1]; ++q) {
				const int i = Li[q];
				if (P.blk_of[i] == P.blk_of[j]) continue;
				P.f_cols[ffill[i]] = j; P.f_vals[ffill[i]] = Lx[q]; ffill[i]++;
				P.b_rows[bfill[j]] = i; P.b_vals[bfill[j]] = Lx[q]; bfill[j]++;
			}
	}

	// ---- levels of the block DAG ----
	std::vector<int> flev(P.n_blocks, 0), blev(P.n_blocks, 0);
	for (int b = 0; b < P.n_blocks; ++b) {
		int l = 0;
		for (int i = start[b]; i < start[b + 1]; ++i) for (int q = P.f_rowptr[i]; q < P.f_rowptr[i + 1]; ++q) l = std::max(l, flev[P.blk_of[P.f_cols[q]]] + 1);
		flev[b] = l; P.n_levels_f = std::max(P.n_levels_f, l + 1);
	}
	for (int b = P.n_blocks - 1; b >= 0; --b) {
		int l = 0;
		for (int j = start[b]; j < start[b + 1]; ++j) for (int q = P.b_colptr[j]; q < P.b_colptr[j + 1]; ++q) l = std::max(l, blev[P.blk_of[P.b_rows[q]]] + 1);
		blev[b] = l; P.n_levels_b = std::max(P.n_levels_b, l + 1);
	}
	auto bucket = [&](const std::vector<int> &lev, int nl, std::vector<int> &ptr, std::vector<int> &items, const std::vector<int> &itemptr, std::vector<int> &lanes) {
		ptr.assign(nl + 1, 0);
		for (int b = 0; b < P.n_blocks; ++b) ptr[lev[b] + 1] += start[b + 1] - start[b];
		for (int l = 0; l < nl; ++l) ptr[l + 1] += ptr[l];
		std::vector<int> fill(ptr.begin(), ptr.end() - 1);
		items.resize(n);
		for (int b = 0; b < P.n_blocks; ++b) for (int i = start[b]; i < start[b + 1]; ++i) items[fill[lev[b]]++] = i;
		lanes.assign(2 * (size_t)nl, 1);
		for (int l = 0; l < nl; ++l) {
			double out = 0, dense = 0;
			const int cnt = ptr[l + 1] - ptr[l];
			for (int k = ptr[l]; k < ptr[l + 1]; ++k) {
				const int i = items[k], b = P.blk_of[i], s = start[b + 1] - start[b];
				out += itemptr[i + 1] - itemptr[i];
				dense += 0.5 * (s - 1);
			}
			lanes[2 * l] = ldlt_pow2_lanes(cnt ? out / cnt : 0.0, cnt);
			lanes[2 * l + 1] = ldlt_pow2_lanes(cnt ? dense / cnt : 0.0, cnt);
		}
	};
	bucket(flev, P.n_levels_f, P.f_lev_ptr, P.f_rows, P.f_rowptr, P.f_lanes);
	bucket(blev, P.n_levels_b, P.b_lev_ptr, P.b_cols, P.b_colptr, P.b_lanes);

	// ---- the bottom forest ----
	// parent block = the block of the elimination-tree parent of a block's last column; a block only reads blocks of its
	// own subtree, so subtrees whose roots sit below the cut are independent of each other
	std::vector<int> bparent(P.n_blocks, -1);
	for (int b = 0; b < P.n_blocks; ++b) {
		const int last = start[b + 1] - 1;
		if (Lp[last + 1] > Lp[last]) bparent[b] = P.blk_of[Li[Lp[last]]];
	}
	// every block a row reads must be a descendant of the row's block: pre-order intervals of the block tree (children have
	// smaller indices than their parent, so one backward sweep sizes the subtrees and one forward... parents first) decide
	bool is_tree = true;
	for (int b = 0; b < P.n_blocks && is_tree; ++b) if (bparent[b] >= 0 && (bparent[b] <= b || flev[bparent[b]] <= flev[b])) is_tree = false;
	if (is_tree) {
		std::vector<int> size(P.n_blocks, 1), tin(P.n_blocks, 0), next_child(P.n_blocks, 0);
		for (int b = 0; b < P.n_blocks; ++b) if (bparent[b] >= 0) size[bparent[b]] += size[b];
		// pre-order number: roots in index order; a child's interval starts inside its parent's
		int counter = 0;
		for (int b = P.n_blocks - 1; b >= 0; --b) {
			if (bparent[b] < 0) { tin[b] = counter; counter += size[b]; next_child[b] = tin[b] + 1; }
			else { tin[b] = next_child[bparent[b]]; next_child[bparent[b]] += size[b]; next_child[b] = tin[b] + 1; }
		}
		for (int i = 0; i < n && is_tree; ++i) {
			const int target = P.blk_of[i];
			for (int q = P.f_rowptr[i]; q < P.f_rowptr[i + 1]; ++q) {
				const int a = P.blk_of[P.f_cols[q]];
				if (tin[a] < tin[target] || tin[a] >= tin[target] + size[target]) { is_tree = false; break; }
			}
		}
	}
	P.n_ctas = n_ctas;
	P.cut = 0;
	if (is_tree) {
		// the highest cut that still leaves at least one subtree per CTA
		for (int cut = P.n_levels_f; cut >= 1; --cut) {
			int roots = 0;
			for (int b = 0; b < P.n_blocks; ++b) if (flev[b] < cut && (bparent[b] < 0 || flev[bparent[b]] >= cut)) ++roots;
			if (roots >= n_ctas || cut == 1) { P.cut = roots >= n_ctas ? cut : 0; break; }
		}
	}
	// row range of (level, block) inside f_rows: blocks were bucketed in ascending order within a level
	std::vector<int> blk_pos(P.n_blocks, 0);
	{
		std::vector<int> fill(P.f_lev_ptr.begin(), P.f_lev_ptr.end() - 1);
		for (int b = 0; b < P.n_blocks; ++b) { blk_pos[b] = fill[flev[b]]; fill[flev[b]] += start[b + 1] - start[b]; }
	}
	P.seg_ptr.assign(n_ctas + 1, 0);
	if (P.cut > 0) {
		// subtree root of every bottom block, work per subtree, longest-processing-time assignment to the CTAs
		std::vector<int> root(P.n_blocks, -1);
		for (int b = P.n_blocks - 1; b >= 0; --b) {
			if (flev[b] >= P.cut) continue;
			root[b] = (bparent[b] >= 0 && flev[bparent[b]] < P.cut) ? root[bparent[b]] : b; // parents have larger indices: already set
		}
		std::vector<double> work(P.n_blocks, 0.0);
		for (int b = 0; b < P.n_blocks; ++b) {
			if (root[b] < 0) continue;
			const double s_ = start[b + 1] - start[b];
			double w = 0.5 * s_ * (s_ - 1) + 8.0 * s_;
			for (int i = start[b]; i < start[b + 1]; ++i) w += P.f_rowptr[i + 1] - P.f_rowptr[i];
			work[root[b]] += w;
		}
		std::vector<int> roots;
		for (int b = 0; b < P.n_blocks; ++b) if (root[b] == b) roots.push_back(b);
		std::sort(roots.begin(), roots.end(), [&](int a, int b) { return work[a] != work[b] ? work[a] > work[b] : a < b; });
		std::vector<double> load(n_ctas, 0.0);
		std::vector<int> cta_of(P.n_blocks, -1);
		for (int r : roots) {
			int best = 0;
			for (int c = 1; c < n_ctas; ++c) if (load[c] < load[best]) best = c;
			cta_of[r] = best; load[best] += work[r];
		}
		// segments: per CTA its subtrees one after the other, inside a subtree level by level; consecutive blocks of one
		// level and one subtree that are adjacent in f_rows merge into one segment
		std::vector<std::vector<int>> blocks_of_cta(n_ctas);
		for (int b = 0; b < P.n_blocks; ++b) if (root[b] >= 0) blocks_of_cta[cta_of[root[b]]].push_back(b);
		for (int c = 0; c < n_ctas; ++c) {
			std::vector<int> &bl = blocks_of_cta[c];
			std::sort(bl.begin(), bl.end(), [&](int a, int b) {
				if (root[a] != root[b]) return root[a] < root[b];
				if (flev[a] != flev[b]) return flev[a] < flev[b];
				return a < b;
			});
			for (size_t k = 0; k < bl.size(); ++k) {
				const int b = bl[k], len = start[b + 1] - start[b];
				const bool extend = !P.seg_begin.empty() && (int)P.seg_begin.size() > P.seg_ptr[c] && k > 0 && root[bl[k - 1]] == root[b] && flev[bl[k - 1]] == flev[b] && P.seg_end.back() == blk_pos[b];
				if (extend) P.seg_end.back() += len;
				else { P.seg_begin.push_back(blk_pos[b]); P.seg_end.push_back(blk_pos[b] + len); P.seg_level.push_back(flev[b]); }
			}
			P.seg_ptr[c + 1] = (int)P.seg_begin.size();
		}
	}
	if (P.seg_begin.empty()) { P.seg_begin.push_back(0); P.seg_end.push_back(0); P.seg_level.push_back(0); }
	// threads per row: inside the forest a row is shared by the threads of ONE CTA, above it by the whole grid
	P.lanes.assign(4 * (size_t)P.n_levels_f, 1);
	for (int l = 0; l < P.n_levels_f; ++l) {
		double fo = 0, bo = 0, dn = 0;
		const int cnt = P.f_lev_ptr[l + 1] - P.f_lev_ptr[l];
		for (int k = P.f_lev_ptr[l]; k < P.f_lev_ptr[l + 1]; ++k) {
			const int i = P.f_rows[k], b = P.blk_of[i], s_ = start[b + 1] - start[b];
			fo += P.f_rowptr[i + 1] - P.f_rowptr[i]; bo += P.b_colptr[i + 1] - P.b_colptr[i]; dn += 0.5 * (s_ - 1);
		}
		const bool bottom = l < P.cut;
		const int threads = bottom ? cta_threads : n_ctas * cta_threads, rows = bottom ? std::max(1, cnt / std::max(1, 2 * n_ctas)) : cnt;
		int t[4] = {ldlt_pow2_lanes(cnt ? fo / cnt : 0.0, rows, threads), ldlt_pow2_lanes(cnt ? dn / cnt : 0.0, rows, threads),
			ldlt_pow2_lanes(cnt ? bo / cnt : 0.0, rows, threads), ldlt_pow2_lanes(cnt ? dn / cnt : 0.0, rows, threads)};
		for (int k = 0; k < 4; ++k) P.lanes[4 * l + k] = bottom ? std::min(t[k], 32) : t[k];
	}
	return P;
}

// Host reference of the block solve (used by the CPU tests): x = P^T L^-T D^-1 L^-1 P b for one right-hand side.
inline void ldlt_blocks_solve_host(const LdltBlockPlan &P, const int *perm, const double *D, const double *b, double *x)
{
	const int n = P.n;
	std::vector<double> t(n), y(n);
	// the device's order: the bottom forest CTA by CTA, segment by segment, then the levels above the cut; backward in
	// reverse, with the SAME (forward) levels -- an ancestor always sits on a higher level than its descendants
	auto fwd = [&](int k0, int k1) {
		for (int k = k0; k < k1; ++k) {
			const int i = P.f_rows[k];
			double s = 0;
			for (int q = P.f_rowptr[i]; q < P.f_rowptr[i + 1]; ++q) s += P.f_vals[q] * y[P.f_cols[q]];
			t[i] = b[perm[i]] - s;
		}
		for (int k = k0; k < k1; ++k) {
			const int i = P.f_rows[k], bl = P.blk_of[i], c0 = P.blk_c0[bl], r = i - c0;
			const double *inv = &P.inv[(size_t)P.inv_off[bl] + (size_t)r * (r - 1) / 2];
			double s = t[i];
			for (int c = 0; c < r; ++c) s += inv[c] * t[c0 + c];
			y[i] = s;
		}
	};
	auto bwd = [&](int k0, int k1) {
		for (int k = k0; k < k1; ++k) {
			const int j = P.f_rows[k];
			double s = 0;
			for (int q = P.b_colptr[j]; q < P.b_colptr[j + 1]; ++q) s += P.b_vals[q] * y[P.b_rows[q]];
			t[j] = y[j] / D[j] - s;
		}
		for (int k = k0; k < k1; ++k) {
			const int j = P.f_rows[k], bl = P.blk_of[j], c0 = P.blk_c0[bl], s_ = P.blk_c0[bl + 1] - c0, r = j - c0;
			const double *invT = &P.invT[(size_t)P.inv_off[bl] + (size_t)r * (s_ - 1) - (size_t)r * (r - 1) / 2];
			double s = t[j];
			for (int c = r + 1; c < s_; ++c) s += invT[c - r - 1] * t[c0 + c];
			y[j] = s;
		}
	};
	if (P.n_ctas > 0 && !P.seg_ptr.empty()) {
		for (int c = 0; c < P.n_ctas; ++c) for (int sg = P.seg_ptr[c]; sg < P.seg_ptr[c + 1]; ++sg) fwd(P.seg_begin[sg], P.seg_end[sg]);
		for (int l = P.cut; l < P.n_levels_f; ++l) fwd(P.f_lev_ptr[l], P.f_lev_ptr[l + 1]);
		for (int l = P.n_levels_f - 1; l >= P.cut; --l) bwd(P.f_lev_ptr[l], P.f_lev_ptr[l + 1]);
		for (int c = 0; c < P.n_ctas; ++c) for (int sg = P.seg_ptr[c + 1] - 1; sg >= P.seg_ptr[c]; --sg) bwd(P.seg_begin[sg], P.seg_end[sg]);
		for (int k = 0; k < n; ++k) x[perm[k]] = y[k];
		return;
	}
	for (int l = 0; l < P.n_levels_f; ++l) {
		for (int k = P.f_lev_ptr[l]; k < P.f_lev_ptr[l + 1]; ++k) {
			const int i = P.f_rows[k];
			double s = 0;
			for (int q = P.f_rowptr[i]; q < P.f_rowptr[i + 1]; ++q) s += P.f_vals[q] * y[P.f_cols[q]];
			t[i] = b[perm[i]] - s;
		}
		for (int k = P.f_lev_ptr[l]; k < P.f_lev_ptr[l + 1]; ++k) {
			const int i = P.f_rows[k], bl = P.blk_of[i], c0 = P.blk_c0[bl], r = i - c0;
			const double *inv = &P.inv[(size_t)P.inv_off[bl] + (size_t)r * (r - 1) / 2];
			double s = t[i];
			for (int c = 0; c < r; ++c) s += inv[c] * t[c0 + c];
			y[i] = s;
		}
	}
	for (int l = 0; l < P.n_levels_b; ++l) {
		for (int k = P.b_lev_ptr[l]; k < P.b_lev_ptr[l + 1]; ++k) {
			const int j = P.b_cols[k];
			double s = 0;
			for (int q = P.b_colptr[j]; q < P.b_colptr[j + 1]; ++q) s += P.b_vals[q] * y[P.b_rows[q]];
			t[j] = y[j] / D[j] - s;
		}
		for (int k = P.b_lev_ptr[l]; k < P.b_lev_ptr[l + 1]; ++k) {
			const int j = P.b_cols[k], bl = P.blk_of[j], c0 = P.blk_c0[bl], s_ = P.blk_c0[bl + 1] - c0, r = j - c0;
			const double *invT = &P.invT[(size_t)P.inv_off[bl] + (size_t)r * (s_ - 1) - (size_t)r * (r - 1) / 2];
			double s = t[j];
			for (int c = r + 1; c < s_; ++c) s += invT[c - r - 1] * t[c0 + c];
			y[j] = s;
		}
	}
	for (int k = 0; k < n; ++k) x[perm[k]] = y[k];
}

} // namespace admmb200
