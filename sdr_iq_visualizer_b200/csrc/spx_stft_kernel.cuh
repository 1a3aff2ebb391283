// spx_stft_kernel.cuh -- K1: the fused STFT kernel and its launcher template.
//
// Persistent CTAs; each CTA hosts FPC "slots" of T = N/16 threads, one frame in flight per slot.
// A slot walks chunks of consecutive frames of one stream, keeps the Welch sum and max-hold of
// its 16 bins per thread in registers, and flushes them with fp64 / ordered-uint atomics at the
// end of every chunk.  Rows (f32 dB, u8, complex) are written straight from the registers of the
// last FFT pass in fftshift order.
#pragma once
#include "spx_stft_device.cuh"
#include "spx_internal.h"

namespace spx {

template <int N>
struct StftCfg {
    static constexpr int T = N / 16;
    static constexpr int FPC = (T >= 256) ? 1 : 256 / T;
    static constexpr int THREADS = T * FPC;
    static constexpr int P = plan_passes(N);
    static constexpr int SLOT_F2 = (P >= 2 ? padded_size(N) : 0) + (P >= 3 ? N : 0);  // float2 per slot
    static constexpr int TW_F2 = plan_tw_size(N);
    static constexpr bool ALL_R16 = (N == 256 || N == 4096 || N == 65536);
};

template <int N>
__device__ __forceinline__ void slot_barrier(int slot) {
    using C = StftCfg<N>;
    if constexpr (C::FPC > 1 && C::T >= 32) {
        asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "n"(C::T) : "memory");
    } else {
        __syncthreads();
    }
}

template <int N, int FMT, bool ACC, int TWM, int OCC>
__global__ void __launch_bounds__(StftCfg<N>::THREADS, OCC) stft_kernel(const StftParams p) {
    using C = StftCfg<N>;
    constexpr int P = C::P;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* smem = reinterpret_cast<float2*>(smem_raw);

    const int slot = threadIdx.x / C::T;
    const int tid = threadIdx.x - slot * C::T;
    float2* bufA = smem + slot * C::SLOT_F2;
    float2* bufB = bufA + padded_size(N);

    const float2* tw = p.tw;
    if constexpr (TWM == TW_SMEM && P > 1) {
        float2* tws = smem + C::FPC * C::SLOT_F2;
        for (int i = threadIdx.x; i < C::TW_F2; i += C::THREADS) tws[i] = __ldg(p.tw + i);
        __syncthreads();
        tw = tws;
    }
    TwRegs<N> twr;
    if constexpr (TWM == TW_REG && P > 1) {
        if constexpr (P > 1) tw_regs_load_pass<N, 1>(twr, tid, p.tw);
        if constexpr (P > 2) tw_regs_load_pass<N, 2>(twr, tid, p.tw);
        if constexpr (P > 3) tw_regs_load_pass<N, 3>(twr, tid, p.tw);
    }

    StftAcc<ACC> acc;
    acc.reset();

    const long long worker = (long long)blockIdx.x * C::FPC + slot;
    const long long n_workers = (long long)gridDim.x * C::FPC;
    const long long iters = (p.total_chunks + n_workers - 1) / n_workers;
    const long long F = p.frames_per_stream;

    float2 v[16];
    for (long long it = 0; it < iters; ++it) {
        const long long chunk = it * n_workers + worker;
        const bool chunk_active = chunk < p.total_chunks;
        const long long stream = chunk_active ? chunk / p.chunks_per_stream : 0;
        const long long f0 = chunk_active ? (chunk - stream * p.chunks_per_stream) * p.frames_per_chunk : 0;
        long long nf = F - f0;
        if (nf > p.frames_per_chunk) nf = p.frames_per_chunk;
        if (!chunk_active) nf = 0;
        const long long sbase = stream * p.stream_stride + f0 * p.hop;
        const long long rbase = stream * F + f0;

        for (int fi = 0; fi < p.frames_per_chunk; ++fi) {
            const bool a = fi < nf;
            const long long s0 = sbase + (long long)fi * p.hop;
            const long long row = rbase + fi;
            stft_phase<N, FMT, ACC, TWM, 0>(v, tid, p, s0, row, a, bufA, bufB, tw, twr, acc);
            if constexpr (P > 1) {
                slot_barrier<N>(slot);
                stft_phase<N, FMT, ACC, TWM, 1>(v, tid, p, s0, row, a, bufA, bufB, tw, twr, acc);
            }
            if constexpr (P > 2) {
                slot_barrier<N>(slot);
                stft_phase<N, FMT, ACC, TWM, 2>(v, tid, p, s0, row, a, bufA, bufB, tw, twr, acc);
            }
            if constexpr (P > 3) {
                slot_barrier<N>(slot);
                stft_phase<N, FMT, ACC, TWM, 3>(v, tid, p, s0, row, a, bufA, bufB, tw, twr, acc);
            }
            // the last pass of an even-P plan reads bufA, which the next frame's pass 0 overwrites
            if constexpr (P > 1 && (P % 2) == 0) slot_barrier<N>(slot);
        }

        if constexpr (ACC) {
            if (chunk_active) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const long long o = stream * N + acc_pos<N>(tid, i);
                    if (p.welch_acc) atomicAdd(p.welch_acc + o, (double)acc.sum[i]);
                    // |X|^2 >= 0: IEEE order == unsigned integer order
                    if (p.maxhold) atomicMax(reinterpret_cast<unsigned int*>(p.maxhold) + o, __float_as_uint(acc.mx[i]));
                }
            }
            acc.reset();
        }
    }
}

// ------------------------------------------------------------------ host side
struct StftLaunch {
    StftParams p;          // chunking fields are filled in by the launcher
    int nfft;
    int in_fmt;
    int variant;           // tuning variant (0 = default)
    int sm_count;
    long long total_frames;  // n_streams * F
    cudaStream_t stream;
};

template <int N, int FMT, bool ACC, int TWM, int OCC>
int launch_stft_inst(StftLaunch& L) {
    using C = StftCfg<N>;
    auto kern = stft_kernel<N, FMT, ACC, TWM, OCC>;
    size_t smem = (size_t)C::FPC * C::SLOT_F2 * sizeof(float2);
    if (TWM == TW_SMEM) smem += (size_t)C::TW_F2 * sizeof(float2);
    static int occ_cache[64] = {0};  // per instantiation, per device (benign race: same value)
    int dev = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    int occ = occ_cache[dev & 63];
    if (occ == 0) {
        SPX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::THREADS, smem));
        if (occ < 1) return spx_set_error(SPX_E_CUDA, "stft kernel does not fit on an SM");
        occ_cache[dev & 63] = occ;
    }
    const long long workers_max = (long long)L.sm_count * occ * C::FPC;
    // chunking: <= 256 frames per chunk, chunks never cross a stream
    const long long F = L.p.frames_per_stream;
    long long per_worker = (L.total_frames + workers_max - 1) / workers_max;
    long long fpc = per_worker < 1 ? 1 : per_worker;
    if (fpc > 256) {
        const long long waves = (per_worker + 255) / 256;
        fpc = (per_worker + waves - 1) / waves;
    }
    if (fpc > F) fpc = F;
    const long long cps = (F + fpc - 1) / fpc;
    L.p.frames_per_chunk = (int)fpc;
    L.p.chunks_per_stream = (int)cps;
    L.p.total_chunks = cps * L.p.n_streams;
    long long grid = (L.p.total_chunks + C::FPC - 1) / C::FPC;
    const long long grid_max = (long long)L.sm_count * occ;
    if (grid > grid_max) grid = grid_max;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, C::THREADS, smem, L.stream>>>(L.p);
    SPX_CUDA(cudaGetLastError());
    return SPX_OK;
}

template <int N, int TWM, int OCC>
int launch_stft_n(StftLaunch& L) {
    const bool acc = L.p.welch_acc != nullptr || L.p.maxhold != nullptr;
    if (L.in_fmt == FMT_CF32) {
        return acc ? launch_stft_inst<N, FMT_CF32, true, TWM, OCC>(L) : launch_stft_inst<N, FMT_CF32, false, TWM, OCC>(L);
    }
    return acc ? launch_stft_inst<N, FMT_CI16, true, TWM, OCC>(L) : launch_stft_inst<N, FMT_CI16, false, TWM, OCC>(L);
}

// one translation unit per size group instantiates these
int launch_stft_small(StftLaunch& L);   // N = 16 .. 512
int launch_stft_1k2k(StftLaunch& L);    // N = 1024, 2048
int launch_stft_4k(StftLaunch& L);      // N = 4096 (+ tuning variants)
int launch_stft_8k(StftLaunch& L);      // N = 8192

}  // namespace spx
