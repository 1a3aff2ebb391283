// Microbenchmark for K4 (I/Q histogram): how many counter increments per clock can one SM retire
//   mode 0: shared-memory atomics on packed 16-bit counters (what hist2d_smem_kernel does)
//   mode 1: global RED into a table PRIVATE to the CTA (no cross-SM contention, table resident in L2)
//   mode 2: alternate samples between the two (does the L2 path run beside the shared-memory atomic unit?)
//   mode 3: 1 sample in 3 to the global table, 2 in 3 to shared memory
//   mode 4: shared-memory atomics with half of the lanes predicated off (is the cost per active lane?)
//   mode 5: global RED into ONE table shared by all CTAs
// dist 0: uniform bins; dist 1: sum of four uniforms per axis (bell-shaped, sigma ~ 21 bins of 256)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/atom_mix.cu -o tools/micro/atom_mix
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s; }

template <int MODE, int DIST>
__global__ void __launch_bounds__(1024) k(unsigned* __restrict__ gpriv, int iters, unsigned* sink) {
    extern __shared__ unsigned sh[];
    for (int w = threadIdx.x; w < 32768; w += blockDim.x) sh[w] = 0u;
    __syncthreads();
    unsigned* g = MODE == 5 ? gpriv : gpriv + (size_t)blockIdx.x * 65536u;
    unsigned s = 12345u + 977u * (blockIdx.x * 1024u + threadIdx.x);
    const bool odd_lane = (threadIdx.x & 1) != 0;
    for (int it = 0; it < iters; ++it) {
        unsigned bi, bq;
        if (DIST == 0) {
            const unsigned r = lcg(s);
            bi = (r >> 24) & 255u;
            bq = (r >> 12) & 255u;
        } else {
            const unsigned a = lcg(s), b = lcg(s);
            bi = (((a >> 26) & 63u) + ((a >> 20) & 63u) + ((a >> 14) & 63u) + ((a >> 8) & 63u)) & 255u;
            bq = (((b >> 26) & 63u) + ((b >> 20) & 63u) + ((b >> 14) & 63u) + ((b >> 8) & 63u)) & 255u;
        }
        const unsigned idx = bi * 256u + bq;
        bool to_global = false, skip = false;
        if (MODE == 1 || MODE == 5) to_global = true;
        if (MODE == 2) to_global = (it & 1) != 0;
        if (MODE == 3) to_global = (it % 3) == 2;
        if (MODE == 4) skip = odd_lane;
        if (skip) continue;
        if (to_global) atomicAdd(g + idx, 1u);
        else atomicAdd(&sh[(idx >> 1) ^ (bi & 31u)], 1u << (16 * (idx & 1u)));
    }
    __syncthreads();
    unsigned acc = 0;
    for (int w = threadIdx.x; w < 32768; w += blockDim.x) acc += sh[w];
    if (acc == 0xdeadbeefu) sink[0] = acc;
}

template <int MODE, int DIST>
static void run(const char* name, unsigned* gpriv, unsigned* sink, int sms, int iters, double mhz) {
    auto kern = k<MODE, DIST>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemsetAsync(gpriv, 0, (size_t)sms * 65536 * 4);
        cudaEventRecord(e0);
        kern<<<sms, 1024, 131072>>>(gpriv, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double ops = (double)sms * 1024.0 * iters * (MODE == 4 ? 0.5 : 1.0);
    printf("{\"mode\": \"%s\", \"dist\": %d, \"ms\": %.4f, \"Gops_per_s\": %.1f, \"ops_per_clk_per_sm\": %.3f, \"us_for_2p24\": %.1f, \"err\": \"%s\"}\n", name,
           DIST, best, ops / best * 1e-6, ops / (best * 1e-3) / (sms * mhz * 1e6), 16777216.0 / (ops / best) * 1e3, cudaGetErrorString(err));
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 256;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    unsigned *gpriv, *sink;
    cudaMalloc(&gpriv, (size_t)sms * 65536 * 4);
    cudaMalloc(&sink, 4);
#define RUN(M, NAME) run<M, 0>(NAME, gpriv, sink, sms, iters, mhz); run<M, 1>(NAME, gpriv, sink, sms, iters, mhz);
    RUN(0, "smem_atoms")
    RUN(1, "global_red_private")
    RUN(2, "alternate_1_1")
    RUN(3, "smem_2_global_1")
    RUN(4, "smem_half_lanes")
    RUN(5, "global_red_shared")
    return 0;
}
