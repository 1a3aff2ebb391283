// Microbenchmark: latency / throughput of a 2-D TMA box load {128 B x 256 rows} (the K1v2 / K2v2 staging shape) as a
// function of the row pitch in global memory and of the number of CTAs per SM, data resident in L2.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/tma_box.cu -o tools/micro/tma_box
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int DEPTH>
__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm, int iters, int x_tiles, int y_tiles, int rows, long long* cycles_out,
                                         unsigned* sink) {
    extern __shared__ unsigned char raw[];
    __shared__ unsigned long long bar[DEPTH];
    unsigned r = s32(raw);
    unsigned char* base = raw + (((r + 1023u) & ~1023u) - r);
    if (threadIdx.x == 0) for (int d = 0; d < DEPTH; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[d])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned acc = 0;
    const int bytes = rows * 128;
    long long t0 = clock64();
    auto issue = [&](int i) {
        const int d = i % DEPTH;
        const int t = (blockIdx.x * 7 + i * 13) % (x_tiles * y_tiles);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[d])), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(base + d * 32768)), "l"(&tm), "r"((t % x_tiles) * 32), "r"((t / x_tiles) * rows), "r"(s32(&bar[d])) : "memory");
    };
    if (threadIdx.x == 0) for (int i = 0; i < DEPTH - 1 && i < iters; ++i) issue(i);
    for (int i = 0; i < iters; ++i) {
        if (threadIdx.x == 0 && i + DEPTH - 1 < iters) issue(i + DEPTH - 1);
        const int d = i % DEPTH;
        const unsigned par = (i / DEPTH) & 1;
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(&bar[d])), "r"(par) : "memory");
        acc += *(unsigned*)(base + d * 32768 + threadIdx.x * 16);
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
}
int main() {
    enc_fn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    const size_t bytes = 64u << 20;
    void* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 1, bytes);
    long long* cyc;
    cudaMalloc(&cyc, 8 * 1024);
    unsigned* sink;
    cudaMalloc(&sink, 4);
    int sm = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 400;
    for (int rows : {256, 64}) for (int pitch : {128, 2048, 8192}) for (int depth : {1, 2}) for (int per_sm : {1, 2}) {
        CUtensorMap tm;
        const cuuint64_t n_rows = bytes / pitch;
        const cuuint64_t gdim[2] = {(cuuint64_t)pitch / 4, n_rows};
        const cuuint64_t gstr[1] = {(cuuint64_t)pitch};
        const cuuint32_t box[2] = {32, (cuuint32_t)rows}, es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        const int x_tiles = pitch / 128, y_tiles = (int)(n_rows / rows);
        const size_t smem = (size_t)depth * 32768 + 1024 + (per_sm == 1 ? 100 * 1024 : 0);   // pad to force 1 CTA/SM
        auto kern = depth == 1 ? k<1> : k<2>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int grid = sm * per_sm;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            kern<<<grid, 256, smem>>>(tm, iters, x_tiles, y_tiles, rows, cyc, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t e = cudaGetLastError();
        printf("{\"rows\": %d, \"pitch\": %d, \"depth\": %d, \"ctas_per_sm\": %d, \"us_per_box_per_cta\": %.3f, \"GBps_total\": %.1f, \"err\": \"%s\"}\n", rows, pitch,
               depth, per_sm, ms * 1e3 / iters, (double)grid * iters * rows * 128 / (ms * 1e-3) / 1e9, cudaGetErrorString(e));
    }
    return 0;
}
