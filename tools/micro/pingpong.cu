// tools/micro/pingpong.cu -- inter-SM signalling latency through L2 on B200 (study aid for the MCGS halo exchange).
// CTA 0 and CTA 1 (different SMs) bounce a tagged 64-bit word: st.relaxed.gpu / ld.relaxed.gpu, as mcgs_*_f32 does.
// Variants: number of words per message (lanes), extra polling threads per SM, all CTAs pairing up.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ void st_ll(uint2 *p, unsigned v, unsigned tag) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(tag) : "memory"); }
__device__ __forceinline__ uint2 ld_ll(const uint2 *p) { uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
// each pair (2i, 2i+1) of CTAs plays ping-pong for `rounds`; message = `lanes` words written by one warp, scattered by `stride` words
__global__ void pingpong(uint2 *buf, int rounds, int lanes, int stride, int pollers, long long *cycles)
{
	const int pair = blockIdx.x >> 1, me = blockIdx.x & 1, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint2 *mine = buf + (size_t)(2 * pair + me) * 4096, *other = buf + (size_t)(2 * pair + (me ^ 1)) * 4096;
	long long t0 = clock64();
	for (int r = 1; r <= rounds; ++r) {
		if (warp == 0) {
			if (me == 0) { // send then wait for the echo
				if (lane < lanes) st_ll(other + lane * stride, r, r);
				if (lane < lanes) while (ld_ll(mine + lane * stride).y != (unsigned)r) { }
			} else {
				if (lane < lanes) while (ld_ll(mine + lane * stride).y != (unsigned)r) { }
				if (lane < lanes) st_ll(other + lane * stride, r, r);
			}
		} else if (warp <= pollers) {
			// extra polling warps spinning on the same round (as the halo pollers do)
			while (ld_ll(mine + (lane % lanes) * stride).y < (unsigned)r) { }
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main(int argc, char **argv)
{
	int n_pairs_list[] = {1, 74};
	uint2 *buf; long long *cyc;
	cudaMalloc(&buf, 148 * 4096 * sizeof(uint2)); cudaMalloc(&cyc, 148 * sizeof(long long));
	const int rounds = 2000;
	for (int np : n_pairs_list) for (int lanes : {1, 32}) for (int stride : {1, 3, 17}) for (int pollers : {0, 4, 15}) {
		cudaMemset(buf, 0, 148 * 4096 * sizeof(uint2));
		pingpong<<<2 * np, 512>>>(buf, rounds, lanes, stride, pollers, cyc);
		cudaError_t e = cudaDeviceSynchronize();
		long long h[148]; cudaMemcpy(h, cyc, sizeof(long long) * 2 * np, cudaMemcpyDeviceToHost);
		double mx = 0; for (int i = 0; i < 2 * np; ++i) mx = h[i] > mx ? h[i] : mx;
		printf("pairs %3d lanes %2d stride %2d pollers %2d: %.0f cycles per one-way hop (%s)\n", np, lanes, stride, pollers, mx / rounds / 2.0, cudaGetErrorString(e));
	}
	return 0;
}
