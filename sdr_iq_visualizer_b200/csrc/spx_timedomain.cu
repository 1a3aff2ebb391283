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

// np.searchsorted(edges, v, 'right') - 1 with the right edge closed; -1 = outside
__device__ __forceinline__ int bin_of(const HistGrid& g, double v) {
    if (!(v >= g.start) || !(v <= g.stop)) return -1;  // also drops NaN
    int b = (int)floor((v - g.start) * g.inv_step);
    b = b < 0 ? 0 : (b > g.bins - 1 ? g.bins - 1 : b);
    while (b > 0 && v < edge_of(g, b)) --b;
    while (b < g.bins - 1 && v >= edge_of(g, b + 1)) ++b;
    return b;
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
        long long blocks = (n + 255) / 256;
        if (blocks > (long long)sm * 8) blocks = (long long)sm * 8;
        const int vec_ok = ((uintptr_t)d_in & 15u) == 0 ? 1 : 0;
        if (in_fmt == SPX_FMT_CF32) hist2d_kernel<SPX_FMT_CF32><<<(unsigned)blocks, 256, 0, st>>>(d_in, n, in_scale, g, d_hist, vec_ok);
        else hist2d_kernel<SPX_FMT_CI16><<<(unsigned)blocks, 256, 0, st>>>(d_in, n, in_scale, g, d_hist, vec_ok);
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
