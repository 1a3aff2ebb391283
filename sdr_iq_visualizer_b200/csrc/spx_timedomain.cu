// spx_timedomain.cu -- K4: time-domain / constellation views.
//
//   * per-frame mean and peak power  mean(I^2+Q^2), max(I^2+Q^2)  (the reference
//     only prints 20*log10(mean|x|) per buffer, /root/reference/scripts/pyad-iio-test.py:93)
//   * 2-D I/Q density histogram with np.histogram2d semantics (A10) replacing the random 2000-point
//     constellation scatter of /root/reference/app/dashboard/callbacks.py:199-214.  Counts are
//     integers and must be bit-exact: the bin of a value is decided against the float64 edge
//     table exactly as numpy builds it (linspace: fl(fl(i*step) + start), last edge = stop).
#include <math.h>
#include <string.h>

#include <mutex>

#include "spx_internal.h"
#include "spx_plan.h"

namespace spx {

struct HistGrid {
    double start, stop, step, inv_step;
    int bins;
};

__device__ __forceinline__ double edge_of(const HistGrid& g, int i) {
    return i >= g.bins ? g.stop : __dadd_rn(__dmul_rn((double)i, g.step), g.start);
}

// np.searchsorted(edges, v, 'right') - 1 with the right edge closed; -1 = outside.
// Branch-light and exact: the float64 estimate floor((v - start) / step) is at most one bin off, so one comparison
// against the lower edge and one against the upper edge -- both built exactly as numpy builds them -- settle it.
__device__ __forceinline__ int bin_of(const HistGrid& g, double v) {
    const double t = (v - g.start) * g.inv_step;
    int b = __double2int_rd(t);
    b = max(0, min(b, g.bins - 1));
    const double e0 = edge_of(g, b);
    b -= (v < e0) ? 1 : 0;
    b = max(b, 0);
    const double e1 = edge_of(g, b + 1);
    b += (v >= e1 && b < g.bins - 1) ? 1 : 0;
    return (v >= g.start && v <= g.stop) ? b : -1;   // NaN fails both comparisons
}

template <int FMT>
__device__ __forceinline__ void load_iq(const void* in, long long i, double scale, double& re, double& im) {
    if (FMT == SPX_FMT_CF32) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(in) + i);
        re = (double)v.x * scale;
        im = (double)v.y * scale;
    } else {
        const short2 v = __ldg(reinterpret_cast<const short2*>(in) + i);
        re = (double)v.x * scale;
        im = (double)v.y * scale;
    }
}

__device__ __forceinline__ void hist_count(const HistGrid& g, unsigned int* hist, double re, double im) {
    const int bi = bin_of(g, re), bq = bin_of(g, im);
    if (bi >= 0 && bq >= 0) atomicAdd(hist + (size_t)bi * g.bins + bq, 1u);  // H[i][j], i <-> I, j <-> Q
}

// Streaming pass over the samples: every thread issues four independent 16-byte loads (2 cf32 or 4 ci16 samples each,
// ld.global.nc, no L1 allocation) before it touches any of them, so 64 bytes per thread are in flight and the pass is
// bandwidth- rather than latency-bound on cold input; the counts go to the L2-resident table with RED.ADD.
// (`vec_ok` = the buffer is 16-byte aligned; otherwise, and for the tail, samples are read one at a time.)
template <int FMT>
__global__ void __launch_bounds__(256) hist2d_kernel(const void* __restrict__ in, long long n, double scale, HistGrid g,
                                                     unsigned int* __restrict__ hist, int vec_ok) {
    constexpr int SPV = FMT == SPX_FMT_CF32 ? 2 : 4;   // samples per 16-byte vector
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nvec = vec_ok ? n / SPV : 0;
    const uint4* vin = reinterpret_cast<const uint4*>(in);
    for (long long v0 = tid0; v0 < nvec; v0 += stride * U) {
        uint4 w[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long vi = v0 + u * stride;
            w[u] = make_uint4(0u, 0u, 0u, 0u);
            if (vi < nvec)
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(w[u].x), "=r"(w[u].y), "=r"(w[u].z), "=r"(w[u].w) : "l"(vin + vi));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v0 + u * stride >= nvec) break;
            const unsigned int q[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
            if (FMT == SPX_FMT_CF32) {
                hist_count(g, hist, (double)__uint_as_float(q[0]) * scale, (double)__uint_as_float(q[1]) * scale);
                hist_count(g, hist, (double)__uint_as_float(q[2]) * scale, (double)__uint_as_float(q[3]) * scale);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    hist_count(g, hist, (double)(short)(q[k] & 0xffffu) * scale, (double)(short)(q[k] >> 16) * scale);
            }
        }
    }
    for (long long i = nvec * SPV + tid0; i < n; i += stride) {   // tail (or everything, when unaligned)
        double re, im;
        load_iq<FMT>(in, i, scale, re, im);
        hist_count(g, hist, re, im);
    }
}

// Shared-memory privatised variant (north_star: "a shared-memory I/Q 2-D histogram for the constellation view"), used
// when bins^2 16-bit counters fit in shared memory (bins <= 320; 128 KB for 256 x 256).  Real constellations pile up
// on a few thousand bins, and RED.ADD on the same L2 sectors from every SM serialises (ncu: 261 us for 2^24 samples with
// the global-atomic kernel above, DRAM at 6 %).  Here a CTA counts a chunk of <= 65 528 samples into packed 16-bit
// counters in shared memory (two per word, ATOMS.ADD of 1 or 1<<16 -- a chunk cannot overflow a 16-bit counter), then
// adds its non-zero counters to the global table: ~8x fewer global atomics, spread out in time.
template <int FMT>
__global__ void __launch_bounds__(1024) hist2d_smem_kernel(const void* __restrict__ in, long long n, double scale, HistGrid g,
                                                           unsigned int* __restrict__ hist, int vec_ok, int chunk) {
    extern __shared__ unsigned int sh[];
    constexpr int SPV = FMT == SPX_FMT_CF32 ? 2 : 4;   // samples per 16-byte vector
    constexpr int U = 4;
    const int bins = g.bins, nb2 = bins * bins, words = (nb2 + 1) / 2;
    // numpy's edge table in shared memory (behind the counters): the exact comparisons become two 8-byte loads instead
    // of int->double conversions and double multiplies per component
    double* edges = reinterpret_cast<double*>(sh + ((words + 1) & ~1));
    for (int i = threadIdx.x; i <= bins; i += blockDim.x) edges[i] = edge_of(g, i);
    const double start = g.start, stop = g.stop;
    const float start_f = (float)g.start, inv_step_f = (float)g.inv_step, scale_f = (float)scale;
    const bool unit_scale = scale == 1.0;
    // the swizzle needs every row to start on a 32-word boundary (bins a multiple of 64); otherwise it is switched off
    const int swz_mask = (bins % 64 == 0) ? 31 : 0;
    // float estimate of the bin (at most one off), settled exactly against the float64 edges
    auto bin_tab = [&](float f, double v) -> int {
        int b = __float2int_rd((f * scale_f - start_f) * inv_step_f);
        b = max(0, min(b, bins - 1));
        b -= (v < edges[b]) ? 1 : 0;
        b = max(b, 0);
        b += (v >= edges[b + 1] && b < bins - 1) ? 1 : 0;
        return (v >= start && v <= stop) ? b : -1;
    };
    // Fast path: when the float estimate t = (v - start)/step sits at least EPS bins away from every edge, floor(t) IS
    // the bin (the float evaluation of t is off by < 1e-4 bins for bins <= 320: three roundings of relative size
    // 2^-24 on magnitudes <= bins/2); only the ~0.4 % of components within EPS of an edge, and everything outside
    // the range, go through the exact float64 comparison against numpy's edge table.
    constexpr float EPS = 2e-3f;
    const float bins_f = (float)bins;
    // nearest integer and distance to it with the 1.5*2^23 magic constant (FADD only, no XU conversions); t < 2^22
    auto split = [&](float t, int& b, bool& sure) {
        const float m = t + 12582912.0f;
        const float rn = m - 12582912.0f;              // rint(t)
        const float d = t - rn;                        // in [-0.5, 0.5]
        b = (__float_as_int(m) - 0x4B400000) - (d < 0.f ? 1 : 0);   // floor(t)
        sure = fabsf(d) > EPS && t > EPS && t < bins_f - EPS;
    };
    auto count_f = [&](float fr, float fi) {
        const float ti = (fr * scale_f - start_f) * inv_step_f, tq = (fi * scale_f - start_f) * inv_step_f;
        int bi, bq;
        bool si, sq;
        split(ti, bi, si);
        split(tq, bq, sq);
        if (!(si && sq)) {
            const double re = unit_scale ? (double)fr : (double)fr * scale, im = unit_scale ? (double)fi : (double)fi * scale;
            bi = bin_tab(fr, re);
            bq = bin_tab(fi, im);
            if (bi < 0 || bq < 0) return;
        }
        // Bank swizzle.  With idx = bi * bins + bq the bank of a counter depends on bq alone (bins/2 words per row is a
        // multiple of 32 for 256 bins), and a constellation cluster is ~15 bins wide: the 32 lanes of a warp fell on ~8
        // banks (ncu: 46.8 % of the shared wavefronts were conflicts).  XOR-ing the word index with the row number spreads
        // a cluster over all banks; it is a bijection inside each aligned group of 32 words, undone by the flush below.
        const int idx = bi * bins + bq;
        atomicAdd(&sh[(idx >> 1) ^ (bi & swz_mask)], 1u << (16 * (idx & 1)));
    };
    const uint4* vin = reinterpret_cast<const uint4*>(in);
    for (long long c = blockIdx.x; c * chunk < n; c += gridDim.x) {
        for (int w = threadIdx.x; w < words; w += blockDim.x) sh[w] = 0u;
        __syncthreads();
        const long long s0 = c * chunk, s1 = (s0 + chunk < n) ? s0 + chunk : n;
        long long done = s0;
        if (vec_ok) {   // chunk is a multiple of 8 samples, so s0 is vector aligned
            const long long v_lo = s0 / SPV, v_hi = s1 / SPV;
            for (long long v0 = v_lo + threadIdx.x; v0 < v_hi; v0 += (long long)blockDim.x * U) {
                uint4 wv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long vi = v0 + (long long)u * blockDim.x;
                    wv[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (vi < v_hi)
                        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(wv[u].x), "=r"(wv[u].y), "=r"(wv[u].z), "=r"(wv[u].w) : "l"(vin + vi));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (v0 + (long long)u * blockDim.x >= v_hi) break;
                    const unsigned int q[4] = {wv[u].x, wv[u].y, wv[u].z, wv[u].w};
                    if (FMT == SPX_FMT_CF32) {
                        count_f(__uint_as_float(q[0]), __uint_as_float(q[1]));
                        count_f(__uint_as_float(q[2]), __uint_as_float(q[3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) count_f((float)(short)(q[k] & 0xffffu), (float)(short)(q[k] >> 16));
                    }
                }
            }
            done = v_hi * SPV;
        }
        for (long long i = done + threadIdx.x; i < s1; i += blockDim.x) {   // tail, or the whole chunk when unaligned
            if (FMT == SPX_FMT_CF32) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(in) + i);
                count_f(v.x, v.y);
            } else {
                const short2 v = __ldg(reinterpret_cast<const short2*>(in) + i);
                count_f((float)v.x, (float)v.y);
            }
        }
        __syncthreads();
        for (int w = threadIdx.x; w < words; w += blockDim.x) {
            const unsigned int v = sh[w];
            const int wo = swz_mask ? (w ^ ((2 * w / bins) & swz_mask)) : w;   // word index before the swizzle (row = 2 w / bins)
            if (v & 0xffffu) atomicAdd(hist + 2 * wo, v & 0xffffu);
            if (v >> 16) atomicAdd(hist + 2 * wo + 1, v >> 16);
        }
        __syncthreads();
    }
}

template <int FMT>
__global__ void __launch_bounds__(256) frame_stats_kernel(const void* __restrict__ in, long long n_frames, int frame_len,
                                                          int hop, float scale, float* __restrict__ mean_pow,
                                                          float* __restrict__ peak_pow) {
    __shared__ double s_sum[8];
    __shared__ float s_max[8];
    for (long long f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const long long base = f * hop;
        double sum = 0.0;
        float mx = 0.f;
        for (int j = threadIdx.x; j < frame_len; j += 256) {
            float re, im;
            if (FMT == SPX_FMT_CF32) {
                const float2 v = __ldg(reinterpret_cast<const float2*>(in) + base + j);
                re = v.x * scale; im = v.y * scale;
            } else {
                const short2 v = __ldg(reinterpret_cast<const short2*>(in) + base + j);
                re = (float)v.x * scale; im = (float)v.y * scale;
            }
            const float p = re * re + im * im;
            sum += (double)p;
            mx = fmaxf(mx, p);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_max[threadIdx.x >> 5] = mx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            float m = 0.f;
            for (int w = 0; w < 8; ++w) { t += s_sum[w]; m = fmaxf(m, s_max[w]); }
            mean_pow[f] = (float)(t / (double)frame_len);
            peak_pow[f] = m;
        }
    }
}

struct TdScratch {
    std::mutex mu;
    DevBuf in, a, b;
    cudaStream_t st = nullptr;
};
static TdScratch g_td[64];

static int td_begin(int device, TdScratch** out) {
    int ndev = 0;
    SPX_TRY(spx_device_count(&ndev));
    if (ndev == 0) return spx_set_error(SPX_E_NODEVICE, "no CUDA device (libspx has no CPU fallback)");
    if (device < 0 || device >= ndev || device >= 64) return spx_set_error(SPX_E_INVALID, "bad device %d", device);
    SPX_CUDA(cudaSetDevice(device));
    *out = &g_td[device];
    return SPX_OK;
}

}  // namespace spx

using namespace spx;

extern "C" int spx_iq_hist2d(int32_t device, int32_t mem, const void* in, int32_t in_fmt, double in_scale, int64_t n,
                             double r, int32_t bins, uint32_t* hist, int32_t accumulate, void* stream) {
    if (!hist) return spx_set_error(SPX_E_INVALID, "hist is NULL");
    if (bins < 1 || bins > 4096 || !(r > 0.0)) return spx_set_error(SPX_E_INVALID, "need 1 <= bins <= 4096 and R > 0");
    if (in_fmt != SPX_FMT_CF32 && in_fmt != SPX_FMT_CI16) return spx_set_error(SPX_E_INVALID, "unknown in_fmt");
    if (n < 0 || (n > 0 && !in)) return spx_set_error(SPX_E_INVALID, "bad input");
    TdScratch* S;
    SPX_TRY(td_begin(device, &S));
    std::lock_guard<std::mutex> lk(S->mu);
    if (!S->st) SPX_CUDA(cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking));
    cudaStream_t st = (mem == SPX_MEM_DEVICE && stream) ? (cudaStream_t)stream : S->st;
    HistGrid g;
    g.start = -r;
    g.stop = r;
    g.step = (g.stop - g.start) / (double)bins;  // numpy linspace: delta / div
    g.inv_step = 1.0 / g.step;
    g.bins = bins;
    const size_t esz = in_fmt == SPX_FMT_CI16 ? 4 : 8, hbytes = (size_t)bins * bins * sizeof(uint32_t);
    const void* d_in = in;
    unsigned int* d_hist = hist;
    if (mem == SPX_MEM_HOST) {
        SPX_TRY(S->in.reserve((size_t)n * esz));
        SPX_TRY(S->a.reserve(hbytes));
        if (n) SPX_CUDA(cudaMemcpyAsync(S->in.ptr, in, (size_t)n * esz, cudaMemcpyHostToDevice, st));
        d_in = S->in.ptr;
        d_hist = (unsigned int*)S->a.ptr;
        if (accumulate) SPX_CUDA(cudaMemcpyAsync(d_hist, hist, hbytes, cudaMemcpyHostToDevice, st));
    }
    if (!accumulate) SPX_CUDA(cudaMemsetAsync(d_hist, 0, hbytes, st));
    if (n > 0) {
        int sm = 148;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
        const int vec_ok = ((uintptr_t)d_in & 15u) == 0 ? 1 : 0;
        const size_t words = (size_t)((bins * bins + 1) / 2);
        const size_t smem = ((words + 1) & ~(size_t)1) * sizeof(unsigned int) + (size_t)(bins + 1) * sizeof(double);
        if (smem <= 200u * 1024u) {
            // shared-memory privatised counting, one CTA per SM, chunks of <= 65 528 samples (16-bit counters)
            const int chunk = 65528;
            auto k_c = hist2d_smem_kernel<SPX_FMT_CF32>;
            auto k_i = hist2d_smem_kernel<SPX_FMT_CI16>;
            SPX_CUDA(cudaFuncSetAttribute(k_c, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            SPX_CUDA(cudaFuncSetAttribute(k_i, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            long long blocks = (n + chunk - 1) / chunk;
            if (blocks > sm) blocks = sm;
            if (in_fmt == SPX_FMT_CF32) k_c<<<(unsigned)blocks, 1024, smem, st>>>(d_in, n, in_scale, g, d_hist, vec_ok, chunk);
            else k_i<<<(unsigned)blocks, 1024, smem, st>>>(d_in, n, in_scale, g, d_hist, vec_ok, chunk);
        } else {
            long long blocks = (n + 255) / 256;
            if (blocks > (long long)sm * 8) blocks = (long long)sm * 8;
            if (in_fmt == SPX_FMT_CF32) hist2d_kernel<SPX_FMT_CF32><<<(unsigned)blocks, 256, 0, st>>>(d_in, n, in_scale, g, d_hist, vec_ok);
            else hist2d_kernel<SPX_FMT_CI16><<<(unsigned)blocks, 256, 0, st>>>(d_in, n, in_scale, g, d_hist, vec_ok);
        }
        SPX_CUDA(cudaGetLastError());
    }
    if (mem == SPX_MEM_HOST) {
        SPX_CUDA(cudaMemcpyAsync(hist, d_hist, hbytes, cudaMemcpyDeviceToHost, st));
        SPX_CUDA(cudaStreamSynchronize(st));
    }
    return SPX_OK;
}

extern "C" int spx_frame_stats(int32_t device, int32_t mem, const void* in, int32_t in_fmt, float in_scale, int64_t n,
                               int32_t frame_len, int32_t hop, float* mean_pow, float* peak_pow, int64_t* n_frames_out,
                               void* stream) {
    if (frame_len < 1 || hop < 1) return spx_set_error(SPX_E_INVALID, "frame_len and hop must be >= 1");
    if (in_fmt != SPX_FMT_CF32 && in_fmt != SPX_FMT_CI16) return spx_set_error(SPX_E_INVALID, "unknown in_fmt");
    const long long F = n < frame_len ? 0 : (n - frame_len) / hop + 1;
    if (n_frames_out) *n_frames_out = F;
    if (F == 0) return SPX_OK;
    if (!in || !mean_pow || !peak_pow) return spx_set_error(SPX_E_INVALID, "NULL buffer");
    TdScratch* S;
    SPX_TRY(td_begin(device, &S));
    std::lock_guard<std::mutex> lk(S->mu);
    if (!S->st) SPX_CUDA(cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking));
    cudaStream_t st = (mem == SPX_MEM_DEVICE && stream) ? (cudaStream_t)stream : S->st;
    const size_t esz = in_fmt == SPX_FMT_CI16 ? 4 : 8;
    const void* d_in = in;
    float *d_mean = mean_pow, *d_peak = peak_pow;
    if (mem == SPX_MEM_HOST) {
        SPX_TRY(S->in.reserve((size_t)n * esz));
        SPX_TRY(S->b.reserve((size_t)F * 2 * sizeof(float)));
        SPX_CUDA(cudaMemcpyAsync(S->in.ptr, in, (size_t)n * esz, cudaMemcpyHostToDevice, st));
        d_in = S->in.ptr;
        d_mean = (float*)S->b.ptr;
        d_peak = d_mean + F;
    }
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    long long blocks = F < (long long)sm * 8 ? F : (long long)sm * 8;
    if (in_fmt == SPX_FMT_CF32)
        frame_stats_kernel<SPX_FMT_CF32><<<(unsigned)blocks, 256, 0, st>>>(d_in, F, frame_len, hop, in_scale, d_mean, d_peak);
    else
        frame_stats_kernel<SPX_FMT_CI16><<<(unsigned)blocks, 256, 0, st>>>(d_in, F, frame_len, hop, in_scale, d_mean, d_peak);
    SPX_CUDA(cudaGetLastError());
    if (mem == SPX_MEM_HOST) {
        SPX_CUDA(cudaMemcpyAsync(mean_pow, d_mean, (size_t)F * sizeof(float), cudaMemcpyDeviceToHost, st));
        SPX_CUDA(cudaMemcpyAsync(peak_pow, d_peak, (size_t)F * sizeof(float), cudaMemcpyDeviceToHost, st));
        SPX_CUDA(cudaStreamSynchronize(st));
    }
    return SPX_OK;
}
