// tools/micro/gather_bench.cu -- what does ONE slice of the resident Gauss-Seidel cost on an SM?  (study aid, not product)
//
// mcgs_owned_f32_kernel spends ~2 200 cycles on a slice (32 nodes x ~18 sliced-ELL rows gathered from shared memory)
// whether 4 or 12 warps are busy; the shared-memory pipe would allow ~200.  This program runs the gather of
// mcgs_owned_f32.cuh on a synthetic part of the bench mesh's size (1 530 own + 820 halo nodes, 873 rows, 48 slices)
// and prints cycles per slice for 1..16 concurrently active warps and for several formulations of the inner loop:
//   v0  owned_gather as shipped (batches of 8, clamped tail)
//   v1  plain loop, unroll 4
//   v2  batches of 8 without the clamp (rows padded to a multiple of 8 in the data)
//   v3  columns pre-multiplied (byte offsets as u16 * 16 via shift hoisted), values as float
//   v4  one 32-bit word per entry: column in the high 16 bits, value as bf16-truncated fp32?  NO -- precision; instead
//       v4 = two entries per 64-bit word (u16 col, u16 col, packed) + separate float2 values: half the index loads
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench tools/micro/gather_bench.cu && ./gather_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void gather_v0(const float *__restrict__ s_val, const uint16_t *__restrict__ s_col, const float4 *__restrict__ s_d,
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

__device__ __forceinline__ void gather_v1(const float *__restrict__ s_val, const uint16_t *__restrict__ s_col, const float4 *__restrict__ s_d,
	int r0, int r1, int lane, float &sx, float &sy, float &sz)
{
	sx = 0.f; sy = 0.f; sz = 0.f;
#pragma unroll 4
	for (int r = r0; r < r1; ++r) {
		const int c = s_col[r * 32 + lane];
		const float a = s_val[r * 32 + lane];
		const float4 dv = s_d[c];
		sx = fmaf(a, dv.x, sx); sy = fmaf(a, dv.y, sy); sz = fmaf(a, dv.z, sz);
	}
}

// whole batches only: the caller guarantees (r1 - r0) % 8 == 0 (padding rows carry a zero coefficient)
__device__ __forceinline__ void gather_v2(const float *__restrict__ s_val, const uint16_t *__restrict__ s_col, const float4 *__restrict__ s_d,
	int r0, int r1, int lane, float &sx, float &sy, float &sz)
{
	sx = 0.f; sy = 0.f; sz = 0.f;
	const float *v = s_val + r0 * 32 + lane;
	const uint16_t *c = s_col + r0 * 32 + lane;
	const int n = r1 - r0;
	for (int r = 0; r < n; r += 8) {
		int cc[8]; float a[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) { cc[j] = c[(r + j) * 32]; a[j] = v[(r + j) * 32]; }
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const float4 dv = s_d[cc[j]];
			sx = fmaf(a[j], dv.x, sx); sy = fmaf(a[j], dv.y, sy); sz = fmaf(a[j], dv.z, sz);
		}
	}
}

// the whole row (<= 24 entries) in flight at once: 3 batches issued before the first use
__device__ __forceinline__ void gather_v3(const float *__restrict__ s_val, const uint16_t *__restrict__ s_col, const float4 *__restrict__ s_d,
	int r0, int r1, int lane, float &sx, float &sy, float &sz)
{
	sx = 0.f; sy = 0.f; sz = 0.f;
	const float *v = s_val + r0 * 32 + lane;
	const uint16_t *c = s_col + r0 * 32 + lane;
	const int n = r1 - r0;
	for (int r = 0; r < n; r += 16) {
		int cc[16]; float a[16];
#pragma unroll
		for (int j = 0; j < 16; ++j) {
			const int rr = min(r + j, n - 1);
			cc[j] = c[rr * 32];
			a[j] = (r + j < n) ? v[rr * 32] : 0.f;
		}
		float4 dv[16];
#pragma unroll
		for (int j = 0; j < 16; ++j) dv[j] = s_d[cc[j]];
#pragma unroll
		for (int j = 0; j < 16; ++j) { sx = fmaf(a[j], dv[j].x, sx); sy = fmaf(a[j], dv[j].y, sy); sz = fmaf(a[j], dv[j].z, sz); }
	}
}

// v4: the columns are stored as BYTE offsets into s_d (col * 16 < 65536 for any part that fits): no multiply between the
// index load and the gather
__device__ __forceinline__ void gather_v4(const float *__restrict__ s_val, const uint16_t *__restrict__ s_off, const float4 *__restrict__ s_d,
	int r0, int r1, int lane, float &sx, float &sy, float &sz)
{
	sx = 0.f; sy = 0.f; sz = 0.f;
	const float *v = s_val + r0 * 32 + lane;
	const uint16_t *c = s_off + r0 * 32 + lane;
	const unsigned char *base = (const unsigned char *)s_d;
	const int n = r1 - r0;
	for (int r = 0; r < n; r += 8) {
		unsigned int cc[8]; float a[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const int rr = min(r + j, n - 1);
			cc[j] = c[rr * 32];
			a[j] = (r + j < n) ? v[rr * 32] : 0.f;
		}
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			const float4 dv = *(const float4 *)(base + cc[j]);
			sx = fmaf(a[j], dv.x, sx); sy = fmaf(a[j], dv.y, sy); sz = fmaf(a[j], dv.z, sz);
		}
	}
}
// v6: two row-steps per index / value load: byte offsets packed in pairs (one LDS.32 = 2 columns), values as float2;
// rows padded to an even length by the caller
__device__ __forceinline__ void gather_v6(const float *__restrict__ s_val, const uint16_t *__restrict__ s_off, const float4 *__restrict__ s_d,
	int r0, int r1, int lane, float &sx, float &sy, float &sz)
{
	sx = 0.f; sy = 0.f; sz = 0.f;
	// pair p of the slice lives at rows (2p, 2p + 1) of the ordinary layout, re-read here as [pair][lane] of 32 / 64 bits
	const float2 *v = (const float2 *)(s_val + r0 * 32) + lane;
	const unsigned int *c = (const unsigned int *)(s_off + r0 * 32) + lane;
	const unsigned char *base = (const unsigned char *)s_d;
	const int np = (r1 - r0) >> 1;
	for (int p = 0; p < np; p += 4) {
		unsigned int cc[4]; float2 a[4];
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int pp = min(p + j, np - 1);
			cc[j] = c[pp * 32];
			a[j] = v[pp * 32];
			if (p + j >= np) a[j] = make_float2(0.f, 0.f);
		}
#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const float4 d0 = *(const float4 *)(base + (cc[j] & 0xffffu)), d1 = *(const float4 *)(base + (cc[j] >> 16));
			sx = fmaf(a[j].x, d0.x, sx); sy = fmaf(a[j].x, d0.y, sy); sz = fmaf(a[j].x, d0.z, sz);
			sx = fmaf(a[j].y, d1.x, sx); sy = fmaf(a[j].y, d1.y, sy); sz = fmaf(a[j].y, d1.z, sz);
		}
	}
}

// v7: the order of the loads is FORCED (volatile asm): all 8 index loads, all 8 value loads, all 8 gathers, then the
// multiply-adds on two independent accumulator sets.  ptxas schedules v0's batch as 2 + 6 (two dependent rounds).
__device__ __forceinline__ unsigned int lds_u16(const void *p) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"((unsigned int)__cvta_generic_to_shared(p))); return v; }
__device__ __forceinline__ float lds_f32(const void *p) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned int)__cvta_generic_to_shared(p))); return v; }
__device__ __forceinline__ float4 lds_f4(unsigned int addr) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v; }
__device__ __forceinline__ void gather_v7(const float *__restrict__ s_val, const uint16_t *__restrict__ s_col, const float4 *__restrict__ s_d,
	int r0, int r1, int lane, float &sx, float &sy, float &sz)
{
	float ax = 0.f, ay = 0.f, az = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
	const float *v = s_val + r0 * 32 + lane;
	const uint16_t *c = s_col + r0 * 32 + lane;
	const unsigned int base = (unsigned int)__cvta_generic_to_shared(s_d);
	const int n = r1 - r0;
	for (int r = 0; r < n; r += 8) {
		unsigned int cc[8]; float a[8]; float4 dv[8];
		int rr[8];
#pragma unroll
		for (int j = 0; j < 8; ++j) rr[j] = min(r + j, n - 1) * 32;
#pragma unroll
		for (int j = 0; j < 8; ++j) cc[j] = lds_u16(c + rr[j]);
#pragma unroll
		for (int j = 0; j < 8; ++j) a[j] = lds_f32(v + rr[j]);
#pragma unroll
		for (int j = 0; j < 8; ++j) dv[j] = lds_f4(base + cc[j] * 16);
#pragma unroll
		for (int j = 0; j < 8; ++j) if (r + j >= n) a[j] = 0.f;
#pragma unroll
		for (int j = 0; j < 8; j += 2) {
			ax = fmaf(a[j], dv[j].x, ax); ay = fmaf(a[j], dv[j].y, ay); az = fmaf(a[j], dv[j].z, az);
			bx = fmaf(a[j + 1], dv[j + 1].x, bx); by = fmaf(a[j + 1], dv[j + 1].y, by); bz = fmaf(a[j + 1], dv[j + 1].z, bz);
		}
	}
	sx = ax + bx; sy = ay + by; sz = az + bz;
}

template <int V>
__global__ void __launch_bounds__(512, 1) bench_kernel(const float *g_val, const uint16_t *g_col, const int *g_srow, int n_loc, int n_rows, int n_slices,
	int active_warps, int reps, int sync_mode, long long *out)
{
	extern __shared__ __align__(128) unsigned char smem[];
	float4 *s_d = (float4 *)smem;
	float *s_val = (float *)(smem + 16 * (size_t)n_loc);
	uint16_t *s_col = (uint16_t *)(smem + 16 * (size_t)n_loc + 4 * 32 * (size_t)n_rows);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int i = tid; i < n_loc; i += blockDim.x) s_d[i] = make_float4(1e-3f * i, 2e-3f * i, -1e-3f * i, 0.f);
	for (int i = tid; i < 32 * n_rows; i += blockDim.x) { s_val[i] = g_val[i]; s_col[i] = g_col[i]; }
	__syncthreads();
	float acc = 0.f;
	long long t0 = clock64();
	for (int rep = 0; rep < reps; ++rep) {
		if (warp < active_warps) {
			const int sl = (warp + rep * 5) % n_slices;
			const int r0 = g_srow[sl], r1 = g_srow[sl + 1];
			float sx, sy, sz;
			if (V == 0) gather_v0(s_val, s_col, s_d, r0, r1, lane, sx, sy, sz);
			else if (V == 1) gather_v1(s_val, s_col, s_d, r0, r1, lane, sx, sy, sz);
			else if (V == 2) gather_v2(s_val, s_col, s_d, r0, r0 + ((r1 - r0) & ~7), lane, sx, sy, sz);
			else if (V == 3) gather_v3(s_val, s_col, s_d, r0, r1, lane, sx, sy, sz);
			else if (V == 4) gather_v4(s_val, s_col, s_d, r0, r1, lane, sx, sy, sz);
			else if (V == 6) gather_v6(s_val, s_col, s_d, r0, r0 + ((r1 - r0) & ~1), lane, sx, sy, sz);
			else gather_v7(s_val, s_col, s_d, r0, r1, lane, sx, sy, sz);
			// the update: write the node's own entry (as the sweep does)
			const int l = (sl * 32 + lane) % n_loc;
			float4 dold = s_d[l];
			s_d[l] = make_float4(0.1f * dold.x + 1e-6f * sx, 0.1f * dold.y + 1e-6f * sy, 0.1f * dold.z + 1e-6f * sz, 0.f);
			acc += sx;
		}
		if (sync_mode == 1) __syncthreads();
	}
	long long t1 = clock64();
	if (lane == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
	if (acc == 123.456f) out[0] = 0;
}

int main()
{
	const int n_own = 1530, n_halo = 820, n_loc = n_own + n_halo, n_slices = 48;
	std::vector<int> srow(n_slices + 1, 0);
	srand(1);
	for (int s = 0; s < n_slices; ++s) srow[s + 1] = srow[s] + 16 + (s % 5); // 16..20 rows per slice, ~873 in total
	const int n_rows = srow[n_slices];
	std::vector<float> val(32 * (size_t)n_rows);
	std::vector<uint16_t> col(32 * (size_t)n_rows);
	for (size_t i = 0; i < val.size(); ++i) { val[i] = (float)(rand() % 1000) * 1e-3f; col[i] = (uint16_t)(rand() % n_loc); }
	// column patterns: what does one LDS.128 gather cost as a function of the bank pattern of its 32 addresses?
	//   random      uniformly random nodes
	//   local       neighbours within +-64 of the node (mesh-like)
	//   linear      col = lane + 32 * row: 32 consecutive float4 = 512 contiguous bytes, conflict-free
	//   samegroup   col = 8 * (lane + row): every lane of a quarter-warp in the same 16-byte bank group (8-way conflict)
	//   broadcast   one word for all lanes
	//   pairs       col = 4 * lane + row: 2 lanes of a quarter-warp per bank group (2-way conflict)
	const char *pat_name[6] = {"random", "local", "linear", "samegroup", "broadcast", "pairs"};
	std::vector<uint16_t> pats[6];
	pats[0] = col;
	for (int p = 1; p < 6; ++p) pats[p].resize(col.size());
	for (int s = 0; s < n_slices; ++s) for (int r = srow[s]; r < srow[s + 1]; ++r) for (int l = 0; l < 32; ++l) {
		const size_t at = (size_t)r * 32 + l;
		int node = (s * 32 + l) % n_own;
		int c = node + (rand() % 129) - 64; if (c < 0) c += n_loc; if (c >= n_loc) c -= n_loc;
		pats[1][at] = (uint16_t)c;
		pats[2][at] = (uint16_t)((l + 32 * r) % n_loc);
		pats[3][at] = (uint16_t)((8 * (l + r)) % (n_loc & ~7));
		pats[4][at] = (uint16_t)(r % n_loc);
		pats[5][at] = (uint16_t)((4 * l + r) % n_loc);
	}
	float *d_val; uint16_t *d_col, *d_pat[6]; int *d_srow; long long *d_out;
	CK(cudaMalloc(&d_val, val.size() * 4)); CK(cudaMalloc(&d_col, col.size() * 2)); for (int p = 0; p < 6; ++p) { CK(cudaMalloc(&d_pat[p], col.size() * 2)); CK(cudaMemcpy(d_pat[p], pats[p].data(), col.size() * 2, cudaMemcpyHostToDevice)); } CK(cudaMalloc(&d_srow, srow.size() * 4)); CK(cudaMalloc(&d_out, 148 * 16 * 8));
	CK(cudaMemcpy(d_val, val.data(), val.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_col, col.data(), col.size() * 2, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_srow, srow.data(), srow.size() * 4, cudaMemcpyHostToDevice));
	const size_t smem = 16 * (size_t)n_loc + 6 * 32 * (size_t)n_rows;
	printf("part: %d local nodes, %d rows, %d slices, %zu B shared memory\n", n_loc, n_rows, n_slices, smem);
	const void *kern[8] = {(const void *)bench_kernel<0>, (const void *)bench_kernel<1>, (const void *)bench_kernel<2>, (const void *)bench_kernel<3>,
		(const void *)bench_kernel<4>, (const void *)bench_kernel<4>, (const void *)bench_kernel<6>, (const void *)bench_kernel<7>};
	for (int v = 0; v < 8; ++v) CK(cudaFuncSetAttribute(kern[v], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	uint16_t *d_off[6];
	for (int p = 0; p < 6; ++p) {
		std::vector<uint16_t> o(pats[p].size());
		for (size_t i = 0; i < o.size(); ++i) o[i] = (uint16_t)(pats[p][i] * 16);
		CK(cudaMalloc(&d_off[p], o.size() * 2)); CK(cudaMemcpy(d_off[p], o.data(), o.size() * 2, cudaMemcpyHostToDevice));
	}
	const int reps = 200;
	for (int pat = 1; pat < 3; ++pat)
		for (int v : {0, 7})
			for (int sync_mode = 0; sync_mode < 1; ++sync_mode)
				for (int aw : {1, 4, 8, 16}) {
					const uint16_t *dc = (v >= 4 && v != 7) ? d_off[pat] : d_pat[pat];
					int nl = n_loc, nr = n_rows, ns = n_slices, r = reps, sm = sync_mode;
					void *args[] = {&d_val, &dc, &d_srow, &nl, &nr, &ns, &aw, &r, &sm, &d_out};
					for (int w = 0; w < 2; ++w) CK(cudaLaunchKernel(kern[v], dim3(148), dim3(512), args, smem, 0));
					CK(cudaDeviceSynchronize());
					long long h[16];
					CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
					printf("columns %-9s v%d %-12s active warps %2d: %6.0f cycles per slice-step (warp 0), %6.0f cycles per slice of pipe time, %5.1f per row-step\n", pat_name[pat], v,
						sync_mode ? "syncthreads" : "free-running", aw, (double)h[0] / reps, (double)h[0] / reps / aw, (double)h[0] / reps / aw / (v == 2 ? 16.0 : 18.0));
				}
	return 0;
}
