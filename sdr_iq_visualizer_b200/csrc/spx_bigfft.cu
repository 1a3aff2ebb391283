// spx_bigfft.cu -- K2: STFT for N = N1*N2 in [16384, 1048576] (BASELINE config 5: N = 65536).
//
// A 65536-point frame is 512 KiB of cf32 and cannot live in one CTA's shared memory, so the
// transform is split four-step style into two fused kernels that meet in an L2-resident scratch
// (frames are processed in batches sized to the 126 MB L2):
//
//   A (columns): for G adjacent columns n2 of the [N1][N2] view of a frame, load
//       x[N2*n1 + n2] * w[N2*n1 + n2] (unpack + window fused; 128-byte row segments), do the N1-point
//       Stockham FFT over n1 in shared memory, multiply by W_N^{n2*k1} (two-level table) and store
//       T[k1][n2] (coalesced along n2).
//   B (rows): for FPC adjacent rows k1 of T, do the N2-point FFT over n2; bin k = k1 + N1*k2.  The
//       fused epilogue (|X|^2, dB, fftshift, Welch, max-hold, u8) is the same as K1's; the dB / u8
//       values are transposed through shared memory so that global stores are runs of FPC bins.
//
// Replaces np.fft.fft for long frames (/root/reference/app/sdr/streamer.py:119; the 2^16-sample
// buffers of /root/reference/scripts/pyad-iio-test.py:50-61).
#include <math.h>
#include <string.h>

#include <vector>

#include "spx_plan.h"
#include "spx_stft_device.cuh"
#include "spx_tables.h"

namespace spx {

struct BigParams {
    const void* in;          // samples of the stream (cf32 / ci16)
    long long sample0;       // first sample of the first frame of this batch
    int hop;
    int frames;              // frames in this batch
    const float* win;        // [N] window * scale or nullptr
    const float2* tw1;       // twiddles of the N1-point FFT
    const float2* tw2;       // twiddles of the N2-point FFT
    const float2* wn_fine;   // W_N^i, i in [0,256)
    const float2* wn_coarse; // W_N^(256 i), i in [0, N/256)
    float2* scratch;         // T: [frames][N1][N2]
    // outputs (rows are indexed from row0)
    long long row0;
    float* db_rows;
    unsigned char* wf_rows;
    float2* spec_rows;
    double* welch_acc;       // [N] of this stream
    float* maxhold;
    float db_eps, db_pw_min, q_vmin, q_scale;
    int frames_per_chunk;    // kernel B: accumulator flush granularity
    int sys_atomics;         // accumulators may live on a peer GPU
};

template <int N1, int N2>
struct BigCfg {
    static constexpr int N = N1 * N2;
    // kernel A
    static constexpr int T1 = N1 / 16;
    static constexpr int G = N1 >= 1024 ? 8 : 16;              // columns per CTA
    static constexpr int THREADS_A = G * T1;
    static constexpr int P1 = plan_passes(N1);
    static constexpr int BUF1 = padded_size(N1) + (P1 >= 3 ? N1 : 0);
    static constexpr int S1 = BUF1 | 1;                        // odd column stride: conflict-free across columns
    // kernel B
    static constexpr int T2 = N2 / 16;
    static constexpr int FPC = N2 >= 1024 ? 8 : (256 / T2 > 32 ? 32 : 256 / T2);  // rows per CTA
    static constexpr int THREADS_B = FPC * T2;
    static constexpr int P2 = plan_passes(N2);
    static constexpr int BUF2 = padded_size(N2) + (P2 >= 3 ? N2 : 0);
    static constexpr int TILE_LD = FPC + 1;                    // padded tile row (floats)
};

// ------------------------------------------------------------------ kernel A: column FFTs + twiddle
template <int N1, int N2, int FMT>
__global__ void __launch_bounds__(BigCfg<N1, N2>::THREADS_A) big_cols_kernel(const BigParams p) {
    using C = BigCfg<N1, N2>;
    constexpr int N = C::N, T1 = C::T1, G = C::G, P = C::P1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* smem = reinterpret_cast<float2*>(smem_raw);
    const int c = threadIdx.x % G;     // column inside the group (fast lane index -> coalesced rows)
    const int tid = threadIdx.x / G;   // position inside the N1-point FFT
    float2* bufA = smem + c * C::S1;
    float2* bufB = bufA + padded_size(N1);
    constexpr int GROUPS = N2 / G;
    const long long items = (long long)p.frames * GROUPS;
    float2 v[16];
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int f = (int)(it / GROUPS);
        const int n2 = (int)(it - (long long)f * GROUPS) * G + c;
        const long long s0 = p.sample0 + (long long)f * p.hop;
        // pass 0 input: x[N2*(tid + t*T1) + n2], window fused
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int i = N2 * (tid + t * T1) + n2;
            if (FMT == FMT_CF32) v[t] = ld_stream_cf32(reinterpret_cast<const float2*>(p.in) + s0 + i);
            else                 v[t] = ld_stream_ci16<TUNE_I2FP>(reinterpret_cast<const short2*>(p.in) + s0 + i);
        }
        if (p.win != nullptr) {
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                const float w = ld_keep(p.win + N2 * (tid + t * T1) + n2);
                v[t].x *= w;
                v[t].y *= w;
            }
        }
        pass_dft<N1, 0>(v);
        if constexpr (P > 1) {
            pass_store_smem<N1, 0>(v, tid, bufA);
            __syncthreads();
            pass_load_smem<N1, 1>(v, tid, bufA);
            pass_twiddle_table<N1, 1, false>(v, tid, p.tw1);
            pass_dft<N1, 1>(v);
        }
        if constexpr (P > 2) {
            pass_store_smem<N1, 1>(v, tid, bufB);
            __syncthreads();
            pass_load_smem<N1, 2>(v, tid, bufB);
            pass_twiddle_table<N1, 2, false>(v, tid, p.tw1);
            pass_dft<N1, 2>(v);
        }
        static_assert(P <= 3, "N1 up to 4096");
        // twiddle W_N^{n2*k1} and store T[k1][n2]
        constexpr int SL = P - 1, RL = plan_radix(N1, SL), NB = 16 / RL;
        float2* trow = p.scratch + (long long)f * N + n2;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
#pragma unroll
            for (int t = 0; t < RL; ++t) {
                const int k1 = tid + T1 * u + t * (N1 / RL);
                const unsigned m = ((unsigned)n2 * (unsigned)k1) & (unsigned)(N - 1);
                const float2 w = cmul(ld_keep(p.wn_coarse + (m >> 8)), ld_keep(p.wn_fine + (m & 255u)));
                trow[(long long)k1 * N2] = cmul(v[u * RL + t], w);
            }
        }
        if constexpr (P > 1) __syncthreads();  // buffers are rewritten by the next item
    }
}

// ------------------------------------------------------------------ kernel B: row FFTs + fused epilogue
template <int N1, int N2, bool ACC>
__global__ void __launch_bounds__(BigCfg<N1, N2>::THREADS_B) big_rows_kernel(const BigParams p) {
    using C = BigCfg<N1, N2>;
    constexpr int N = C::N, T2 = C::T2, FPC = C::FPC, P = C::P2, LD = C::TILE_LD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* smem = reinterpret_cast<float2*>(smem_raw);
    float* tile = reinterpret_cast<float*>(smem_raw);            // aliases the exchange buffers (after a barrier)
    const int slot = threadIdx.x / T2;
    const int tid = threadIdx.x - slot * T2;
    float2* bufA = smem + slot * C::BUF2;
    float2* bufB = bufA + padded_size(N2);
    constexpr int GROUPS = N1 / FPC;
    const int chunks = (p.frames + p.frames_per_chunk - 1) / p.frames_per_chunk;
    const long long items = (long long)GROUPS * chunks;
    constexpr int SL = P - 1, RL = plan_radix(N2, SL), NB = 16 / RL;
    StftAcc<ACC> acc;
    acc.reset();
    float2 v[16];
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int g = (int)(it % GROUPS);        // consecutive CTAs take consecutive row groups of the same frames
        const int ch = (int)(it / GROUPS);
        const int k1 = g * FPC + slot;
        const int f_lo = ch * p.frames_per_chunk;
        const int f_hi = min(p.frames, f_lo + p.frames_per_chunk);
        for (int f = f_lo; f < f_hi; ++f) {
            const float2* row = p.scratch + (long long)f * N + (long long)k1 * N2;
#pragma unroll
            for (int t = 0; t < 16; ++t) v[t] = row[tid + t * T2];
            pass_dft<N2, 0>(v);
            if constexpr (P > 1) {
                pass_store_smem<N2, 0>(v, tid, bufA);
                __syncthreads();
                pass_load_smem<N2, 1>(v, tid, bufA);
                pass_twiddle_table<N2, 1, false>(v, tid, p.tw2);
                pass_dft<N2, 1>(v);
            }
            if constexpr (P > 2) {
                pass_store_smem<N2, 1>(v, tid, bufB);
                __syncthreads();
                pass_load_smem<N2, 2>(v, tid, bufB);
                pass_twiddle_table<N2, 2, false>(v, tid, p.tw2);
                pass_dft<N2, 2>(v);
            }
            static_assert(P <= 3, "N2 up to 4096");
            const long long orow = (p.row0 + f) * (long long)N;
            // bin k = k1 + N1*k2 -> fftshift position k1 + N1*((k2 + N2/2) mod N2)
            if (p.spec_rows) {
#pragma unroll
                for (int u = 0; u < NB; ++u)
#pragma unroll
                    for (int t = 0; t < RL; ++t) {
                        const int k2s = (tid + T2 * u + t * (N2 / RL) + N2 / 2) & (N2 - 1);
                        p.spec_rows[orow + k1 + (long long)N1 * k2s] = v[u * RL + t];
                    }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i].x = v[i].x * v[i].x + v[i].y * v[i].y;
            if (ACC) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    acc.sum[i] += v[i].x;
                    acc.mx[i] = fmaxf(acc.mx[i], v[i].x);
                }
            }
            if (p.db_rows || p.wf_rows) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    v[i].y = v[i].x >= p.db_pw_min ? amp_db_fast(v[i].x) : amp_db_exact(v[i].x, p.db_eps);
                __syncthreads();  // every thread is done reading the exchange buffers: reuse them as the tile
#pragma unroll
                for (int u = 0; u < NB; ++u)
#pragma unroll
                    for (int t = 0; t < RL; ++t) {
                        const int k2s = (tid + T2 * u + t * (N2 / RL) + N2 / 2) & (N2 - 1);
                        tile[k2s * LD + slot] = v[u * RL + t].y;
                    }
                __syncthreads();
                // copy out: runs of FPC consecutive bins k1_0 .. k1_0+FPC-1 for every k2
                const int k1_0 = g * FPC;
                for (int idx = threadIdx.x; idx < FPC * N2; idx += C::THREADS_B) {
                    const int j = idx % FPC, k2s = idx / FPC;
                    const float db = tile[k2s * LD + j];
                    const long long o = orow + k1_0 + j + (long long)N1 * k2s;
                    if (p.db_rows) p.db_rows[o] = db;
                    if (p.wf_rows) p.wf_rows[o] = (unsigned char)sat_floor_u8((db - p.q_vmin) * p.q_scale);
                }
                __syncthreads();  // tile is overwritten by the next frame's pass 0
            } else if constexpr (P > 1) {
                __syncthreads();
            }
        }
        if constexpr (ACC) {
#pragma unroll
            for (int u = 0; u < NB; ++u)
#pragma unroll
                for (int t = 0; t < RL; ++t) {
                    const int k2s = (tid + T2 * u + t * (N2 / RL) + N2 / 2) & (N2 - 1);
                    const long long o = k1 + (long long)N1 * k2s;
                    flush_acc(p.welch_acc, p.maxhold, o, acc.sum[u * RL + t], acc.mx[u * RL + t], p.sys_atomics);
                }
            acc.reset();
        }
    }
}

// ------------------------------------------------------------------ host side
static void split_n(int n, int* n1, int* n2) {
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    *n1 = 1 << (lg / 2);
    *n2 = n / *n1;
}

int bigfft_plan_init(spx_plan* pl) {
    const int n = pl->cfg.nfft;
    int n1, n2;
    split_n(n, &n1, &n2);
    pl->big_n1 = n1;
    pl->big_n2 = n2;
    std::vector<float2> t1 = build_twiddles(n1), t2 = build_twiddles(n2);
    std::vector<float2> fine(256), coarse((size_t)n / 256);
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < 256; ++i) {
        const double a = -two_pi * (double)i / (double)n;
        fine[i] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int i = 0; i < n / 256; ++i) {
        const double a = -two_pi * (double)i * 256.0 / (double)n;
        coarse[i] = make_float2((float)cos(a), (float)sin(a));
    }
    const size_t total = t1.size() + t2.size() + fine.size() + coarse.size();
    SPX_CUDA(cudaMalloc(&pl->d_big_tw, total * sizeof(float2)));
    float2* d = pl->d_big_tw;
    pl->d_tw1 = d;
    SPX_CUDA(cudaMemcpy(d, t1.data(), t1.size() * sizeof(float2), cudaMemcpyHostToDevice));
    d += t1.size();
    pl->d_tw2 = d;
    SPX_CUDA(cudaMemcpy(d, t2.data(), t2.size() * sizeof(float2), cudaMemcpyHostToDevice));
    d += t2.size();
    pl->d_wn_fine = d;
    SPX_CUDA(cudaMemcpy(d, fine.data(), fine.size() * sizeof(float2), cudaMemcpyHostToDevice));
    d += fine.size();
    pl->d_wn_coarse = d;
    SPX_CUDA(cudaMemcpy(d, coarse.data(), coarse.size() * sizeof(float2), cudaMemcpyHostToDevice));
    return SPX_OK;
}

template <int N1, int N2>
static int big_launch_pair(spx_plan* pl, BigParams& p, cudaStream_t st) {
    using C = BigCfg<N1, N2>;
    const bool acc = p.welch_acc != nullptr || p.maxhold != nullptr;
    const size_t smem_a = (size_t)C::G * C::S1 * sizeof(float2);
    const size_t smem_b_buf = (size_t)C::FPC * C::BUF2 * sizeof(float2);
    const size_t smem_b_tile = (size_t)N2 * C::TILE_LD * sizeof(float);
    const size_t smem_b = smem_b_buf > smem_b_tile ? smem_b_buf : smem_b_tile;
    auto ka_c = big_cols_kernel<N1, N2, FMT_CF32>;
    auto ka_i = big_cols_kernel<N1, N2, FMT_CI16>;
    auto kb_a = big_rows_kernel<N1, N2, true>;
    auto kb_n = big_rows_kernel<N1, N2, false>;
    static bool configured[64] = {false};
    int dev = 0;
    SPX_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        SPX_CUDA(cudaFuncSetAttribute(ka_c, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
        SPX_CUDA(cudaFuncSetAttribute(ka_i, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
        SPX_CUDA(cudaFuncSetAttribute(kb_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        SPX_CUDA(cudaFuncSetAttribute(kb_n, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        configured[dev & 63] = true;
    }
    const long long items_a = (long long)p.frames * (N2 / C::G);
    long long grid_a = items_a < (long long)pl->sm_count * 4 ? items_a : (long long)pl->sm_count * 4;
    if (pl->cfg.in_fmt == SPX_FMT_CF32) ka_c<<<(unsigned)grid_a, C::THREADS_A, smem_a, st>>>(p);
    else ka_i<<<(unsigned)grid_a, C::THREADS_A, smem_a, st>>>(p);
    SPX_CUDA(cudaGetLastError());
    // kernel B: (row groups) x (frame chunks) work items; chunk size balances SM fill vs atomic flushes
    const int groups = N1 / C::FPC;
    const int want_items = pl->sm_count * 2;
    int chunks = (want_items + groups - 1) / groups;
    if (chunks > p.frames) chunks = p.frames;
    if (chunks < 1) chunks = 1;
    p.frames_per_chunk = (p.frames + chunks - 1) / chunks;
    const long long items_b = (long long)groups * ((p.frames + p.frames_per_chunk - 1) / p.frames_per_chunk);
    long long grid_b = items_b < (long long)pl->sm_count * 4 ? items_b : (long long)pl->sm_count * 4;
    if (acc) kb_a<<<(unsigned)grid_b, C::THREADS_B, smem_b, st>>>(p);
    else kb_n<<<(unsigned)grid_b, C::THREADS_B, smem_b, st>>>(p);
    SPX_CUDA(cudaGetLastError());
    return SPX_OK;
}

// one stream, frames [0, frames): batches of frames through the L2-sized scratch
int bigfft_launch_stream(spx_plan* pl, const void* in, long long frames, long long row0, float* db_rows,
                         unsigned char* wf_rows, float2* spec_rows, double* welch_acc, float* maxhold, float vmin,
                         float vmax, cudaStream_t st, int sys_atomics) {
    const int n = pl->cfg.nfft;
    const size_t frame_bytes = (size_t)n * sizeof(float2);
    long long fb = (long long)(pl->big_scratch_bytes / frame_bytes);
    if (fb < 1) fb = 1;
    if (fb > frames) fb = frames;
    SPX_TRY(pl->st_big.reserve((size_t)fb * frame_bytes));
    BigParams p;
    memset(&p, 0, sizeof(p));
    p.in = in;
    p.hop = pl->cfg.hop;
    p.win = pl->d_win;
    p.tw1 = pl->d_tw1;
    p.tw2 = pl->d_tw2;
    p.wn_fine = pl->d_wn_fine;
    p.wn_coarse = pl->d_wn_coarse;
    p.scratch = (float2*)pl->st_big.ptr;
    p.db_rows = db_rows;
    p.wf_rows = wf_rows;
    p.spec_rows = spec_rows;
    p.welch_acc = welch_acc;
    p.maxhold = maxhold;
    p.db_eps = pl->cfg.db_eps;
    p.db_pw_min = pl->cfg.db_eps * pl->cfg.db_eps * 1099511627776.0f;
    p.q_vmin = vmin;
    p.q_scale = 256.0f / (vmax - vmin);
    p.sys_atomics = sys_atomics;
    for (long long f0 = 0; f0 < frames; f0 += fb) {
        p.frames = (int)(frames - f0 < fb ? frames - f0 : fb);
        p.sample0 = f0 * pl->cfg.hop;
        p.row0 = row0 + f0;
        int rc;
        switch (n) {
            case 1 << 14: rc = big_launch_pair<128, 128>(pl, p, st); break;
            case 1 << 15: rc = big_launch_pair<128, 256>(pl, p, st); break;
            case 1 << 16: rc = big_launch_pair<256, 256>(pl, p, st); break;
            case 1 << 17: rc = big_launch_pair<256, 512>(pl, p, st); break;
            case 1 << 18: rc = big_launch_pair<512, 512>(pl, p, st); break;
            case 1 << 19: rc = big_launch_pair<512, 1024>(pl, p, st); break;
            case 1 << 20: rc = big_launch_pair<1024, 1024>(pl, p, st); break;
            default: return spx_set_error(SPX_E_UNSUPPORTED, "nfft %d", n);
        }
        SPX_TRY(rc);
    }
    return SPX_OK;
}

}  // namespace spx
