// store_bench.cu -- what does PUBLISHING cost?  One warp of a CTA stores 12 words per lane per round (3 components x 4 readers,
// the static-ownership solve kernel's worst case) into mailbox-like arrays; cycles per round by store flavour and address pattern.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o _build/store_bench store_bench.cu && ./_build/store_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void st_relaxed_v2(uint2 *p, unsigned a, unsigned b) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void st_weak_v2(uint2 *p, unsigned a, unsigned b) { asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void st_relaxed_v4(uint4 *p, unsigned a, unsigned b, unsigned c, unsigned d) { asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory"); }
__device__ __forceinline__ void st_weak_v4(uint4 *p, unsigned a, unsigned b, unsigned c, unsigned d) { asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory"); }

// mode 0: 12 x relaxed v2 (3 components TS apart, 4 readers); 1: the same with weak stores; 2: 4 x relaxed v4 (one 16-byte word per reader);
// 3: 4 x weak v4.  pattern 0: lane -> consecutive slots of a reader; 1: lane -> random slot
__global__ void bench(uint2 *buf, size_t TS, int mode, int pattern, int active_warps, int rounds, long long *out, const int *rnd)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (warp >= active_warps) return;
	long long t0 = 0, t1 = 0;
	size_t base = ((size_t)blockIdx.x * 32 + warp) * 4096;
	for (int rep = 0; rep < 2; ++rep) {
		__syncwarp();
		t0 = clock64();
		for (int r = 0; r < rounds; ++r) {
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const size_t slot = base + (size_t)q * 1024 + (pattern ? (size_t)rnd[(r * 4 + q) * 32 + lane] : (size_t)((r * 32 + lane) & 1023));
				if (mode == 0) { st_relaxed_v2(buf + slot, r, r + 1); st_relaxed_v2(buf + TS + slot, r, r + 1); st_relaxed_v2(buf + 2 * TS + slot, r, r + 1); }
				else if (mode == 1) { st_weak_v2(buf + slot, r, r + 1); st_weak_v2(buf + TS + slot, r, r + 1); st_weak_v2(buf + 2 * TS + slot, r, r + 1); }
				else if (mode == 2) st_relaxed_v4((uint4 *)buf + slot, r, r + 1, r + 2, r + 3);
				else st_weak_v4((uint4 *)buf + slot, r, r + 1, r + 2, r + 3);
			}
		}
		t1 = clock64();
	}
	if (lane == 0 && warp == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / rounds;
}

int main()
{
	const size_t TS = 148 * 32 * 4096;
	uint2 *buf; long long *out; int *rnd;
	CK(cudaMalloc(&buf, 2 * 3 * TS * sizeof(uint2)));
	CK(cudaMalloc(&out, 64));
	const int rounds = 64;
	int *h = new int[rounds * 4 * 32];
	for (int i = 0; i < rounds * 4 * 32; ++i) h[i] = (int)((1103515245u * (unsigned)i + 12345u) >> 8) & 1023;
	CK(cudaMalloc(&rnd, rounds * 4 * 32 * sizeof(int)));
	CK(cudaMemcpy(rnd, h, rounds * 4 * 32 * sizeof(int), cudaMemcpyHostToDevice));
	const char *mn[4] = {"12 x st.relaxed.gpu.v2", "12 x st.weak.v2       ", " 4 x st.relaxed.gpu.v4", " 4 x st.weak.v4       "};
	for (int grid : {1, 148})
		for (int pattern = 0; pattern < 2; ++pattern)
			for (int aw : {1, 4})
				for (int mode = 0; mode < 4; ++mode) {
					bench<<<grid, 512>>>(buf, TS, mode, pattern, aw, rounds, out, rnd);
					CK(cudaDeviceSynchronize());
					long long c; CK(cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost));
					printf("CTAs %3d  %s slots  warps/CTA %d  %s : %5lld cycles per round (32 lanes x 4 readers x 3 components)\n", grid, pattern ? "random     " : "consecutive", aw, mn[mode], c);
				}
	return 0;
}
