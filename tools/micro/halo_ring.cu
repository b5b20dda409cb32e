// tools/micro/halo_ring.cu -- the communication skeleton of the resident Gauss-Seidel (mcgs_owned_f32.cuh)
// without the arithmetic: 148 persistent CTAs, each with a fixed set of neighbours; per "pass" every CTA waits for
// the tagged words of ALL its neighbours (previous pass), spends `delay` cycles (the boundary gather), publishes
// its own words to every neighbour and ends the pass with a CTA barrier.  Prints cycles per pass for
//   neighbours per CTA (2 .. 26)  x  words per neighbour  x  delay  x  polling strategy,
// i.e. how much of the ~5 500 cycles per colour pass is the price of waiting for the slowest of
// N neighbours 120 times in a row, and which protocol variant lowers it.  Study aid; nothing links against it.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/halo_ring tools/micro/halo_ring.cu && timeout 120 /tmp/halo_ring
// (NOT RUN YET: written after this round's GPU budget was spent; run it under `timeout`, it spins on device memory.)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_ll(uint2 *p, unsigned v, unsigned tag) { asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(tag) : "memory"); }
__device__ __forceinline__ uint2 ld_ll(const uint2 *p) { uint2 v; asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acq(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct Params {
	int n_nbr;            // neighbours per CTA
	int words;            // 64-bit {value, tag} words per directed edge
	int passes;
	int delay;            // cycles of "boundary gather" between arrival and publish
	int strategy;         // 0: poll every word (flag in data); 1: one sentinel lane per warp first, then every word;
	                      // 2: data words + ONE flag word per edge (st.release after the data, ld.acquire polls), data read once
	int n_poll_warps;     // warps that poll (the rest idles at the end-of-pass barrier, like warps with interior slices)
	const int *nbr;       // [grid][n_nbr] neighbour CTA ids
	const int *slot;      // [grid][n_nbr] my index in that neighbour's incoming table
	uint2 *mbox;          // [2][grid][n_nbr][words]  incoming, double-buffered by pass parity
	unsigned *flag;       // [2][grid][n_nbr]
	long long *cycles;    // [grid]
};

__global__ void __launch_bounds__(512, 1) halo_ring(Params P)
{
	extern __shared__ float sink[];
	const int me = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const size_t stride = (size_t)gridDim.x * P.n_nbr * P.words, fstride = (size_t)gridDim.x * P.n_nbr;
	const int n_in = P.n_nbr * P.words;                   // words this CTA waits for per pass
	const int n_poll = 32 * P.n_poll_warps;
	const long long t0 = clock64();
	for (int p = 0; p < P.passes; ++p) {
		const unsigned tag = (unsigned)p;                  // what the neighbours published in pass p - 1
		if (warp < P.n_poll_warps) {
			if (p > 0) {
				const uint2 *in = P.mbox + (size_t)((p - 1) & 1) * stride + (size_t)me * n_in;
				if (P.strategy == 2) {
					const unsigned *f = P.flag + (size_t)((p - 1) & 1) * fstride + (size_t)me * P.n_nbr;
					for (int e = tid; e < P.n_nbr; e += n_poll) while (ld_acq(f + e) != tag) { }
					named_sync(2, n_poll);
					for (int w = tid; w < n_in; w += n_poll) sink[w & 1023] = __uint_as_float(ld_ll(in + w).x);
				} else {
					for (int base = 0; base < n_in; base += n_poll) { // the same trip count for every lane (__syncwarp inside)
						const int w = base + tid;
						const bool valid = w < n_in;
						if (P.strategy == 1) { if (valid && (lane & 15) == 0) while (ld_ll(in + w).y != tag) { } __syncwarp(); }
						if (valid) {
							uint2 v = ld_ll(in + w);
							while (v.y != tag) v = ld_ll(in + w);
							sink[w & 1023] = __uint_as_float(v.x);
						}
					}
				}
			}
			named_sync(1, n_poll);
			// the "boundary gather"
			const long long t = clock64();
			while (clock64() - t < P.delay) { }
			// publish: word w of edge e goes to neighbour nbr[e], into its table at my slot
			for (int w = tid; w < n_in; w += n_poll) {
				const int e = w / P.words, k = w % P.words;
				const int dst = P.nbr[me * P.n_nbr + e], s = P.slot[me * P.n_nbr + e];
				st_ll(P.mbox + (size_t)(p & 1) * stride + ((size_t)dst * P.n_nbr + s) * P.words + k, (unsigned)(me + p), (unsigned)(p + 1));
			}
			if (P.strategy == 2) {
				named_sync(2, n_poll);                     // every data word of this CTA has been issued
				for (int e = tid; e < P.n_nbr; e += n_poll) {
					const int dst = P.nbr[me * P.n_nbr + e], s = P.slot[me * P.n_nbr + e];
					__threadfence();
					st_rel(P.flag + (size_t)(p & 1) * fstride + (size_t)dst * P.n_nbr + s, (unsigned)(p + 1));
				}
			}
		}
		__syncthreads();
	}
	if (tid == 0) P.cycles[me] = clock64() - t0;
}

int main()
{
	int dev = 0, sms = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const int grid = sms;
	cudaFuncSetAttribute(halo_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); // one CTA per SM, like the solve kernel
	printf("%d CTAs x 512 threads, 240 passes; cycles per pass (max over CTAs)\n", grid);
	printf("%-4s %-6s %-6s %-9s %-6s %s\n", "nbr", "words", "delay", "strategy", "pollw", "cycles/pass");
	for (int n_nbr : {2, 6, 12, 26}) {
		// ring of CTAs: neighbours c +- 1, +- 2, ... (edge e and e ^ 1 mirror each other)
		std::vector<int> nbr((size_t)grid * n_nbr), slot((size_t)grid * n_nbr);
		for (int c = 0; c < grid; ++c) for (int e = 0; e < n_nbr; ++e) {
			const int k = e / 2 + 1, sgn = e % 2 ? -1 : 1;
			nbr[(size_t)c * n_nbr + e] = ((c + sgn * k) % grid + grid) % grid;
			slot[(size_t)c * n_nbr + e] = e ^ 1;
		}
		int *d_nbr, *d_slot; long long *d_cyc;
		cudaMalloc(&d_nbr, nbr.size() * 4); cudaMalloc(&d_slot, slot.size() * 4); cudaMalloc(&d_cyc, grid * 8);
		cudaMemcpy(d_nbr, nbr.data(), nbr.size() * 4, cudaMemcpyHostToDevice);
		cudaMemcpy(d_slot, slot.data(), slot.size() * 4, cudaMemcpyHostToDevice);
		for (int words : {12, 48}) for (int delay : {0, 800, 1600}) for (int strategy : {0, 1, 2}) for (int pollw : {4, 11}) {
			Params P;
			P.n_nbr = n_nbr; P.words = words; P.passes = 240; P.delay = delay; P.strategy = strategy; P.n_poll_warps = pollw;
			P.nbr = d_nbr; P.slot = d_slot; P.cycles = d_cyc;
			const size_t nw = (size_t)2 * grid * n_nbr * words;
			cudaMalloc(&P.mbox, nw * sizeof(uint2)); cudaMemset(P.mbox, 0xff, nw * sizeof(uint2));
			cudaMalloc(&P.flag, (size_t)2 * grid * n_nbr * 4); cudaMemset(P.flag, 0xff, (size_t)2 * grid * n_nbr * 4);
			void *args[] = {&P};
			cudaLaunchCooperativeKernel((void *)halo_ring, dim3(grid), dim3(512), args, 200 * 1024, 0);
			cudaError_t e = cudaDeviceSynchronize();
			std::vector<long long> h(grid);
			cudaMemcpy(h.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
			const long long mx = *std::max_element(h.begin(), h.end());
			printf("%-4d %-6d %-6d %-9d %-6d %8.0f   %s\n", n_nbr, words, delay, strategy, pollw, (double)mx / P.passes, e == cudaSuccess ? "" : cudaGetErrorString(e));
			cudaFree(P.mbox); cudaFree(P.flag);
		}
		cudaFree(d_nbr); cudaFree(d_slot); cudaFree(d_cyc);
	}
	return 0;
}
