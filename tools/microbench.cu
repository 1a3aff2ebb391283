// microbench.cu -- FP32 issue-rate probes for sm_100a: scalar FFMA/FADD vs packed FFMA2/FADD2.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Output: one JSON line per probe with achieved TFLOP/s (FMA = 2 flop, ADD = 1 flop per lane-op).
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float2 r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-4f - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { r[i].x = fmaf(r[i].x, a, b); r[i].y = fmaf(r[i].y, a, b); }       // 2 FFMA
            if (MODE == 1) { r[i] = __ffma2_rn(r[i], a2, b2); }                                  // 1 FFMA2
            if (MODE == 2) { r[i].x = r[i].x + b; r[i].y = r[i].y + a; }                         // 2 FADD
            if (MODE == 3) { r[i] = __fadd2_rn(r[i], b2); }                                      // 1 FADD2
            if (MODE == 4) { r[i].x = fmaf(r[i].x, a, b); r[i].y = r[i].y + a; }                 // FFMA + FADD mix
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double flop_per_pair) {
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sm * 8, iters = 1 << 14;
    float* out;
    cudaMalloc(&out, sizeof(float) * blocks * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int k = 0; k < 5; ++k) {
        cudaEventRecord(e0);
        probe<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double pairs = (double)blocks * 256 * iters * 8;  // float2 element updates
    printf("{\"probe\": \"%s\", \"ms\": %.4f, \"Tpairs_per_s\": %.3f, \"TFLOPs\": %.2f}\n", name, best,
           pairs / (best * 1e-3) / 1e12, pairs * flop_per_pair / (best * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    run<0>("2xFFMA scalar", 4.0);
    run<1>("FFMA2 packed", 4.0);
    run<2>("2xFADD scalar", 2.0);
    run<3>("FADD2 packed", 2.0);
    run<4>("FFMA+FADD mix", 3.0);
    return 0;
}
