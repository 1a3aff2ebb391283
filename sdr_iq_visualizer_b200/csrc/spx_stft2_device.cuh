// spx_stft2_device.cuh -- per-thread phases of K1v2, the "warp-local first exchange" fused STFT kernel for
// N = 256*C, C in {4, 8, 16} (N = 1024, 2048, 4096).
//
// Same job as K1 (spx_stft_device.cuh): unpack -> window -> FFT -> |X|^2 -> dB -> fftshift -> {f32 rows, u8 rows, Welch,
// max-hold}, replacing /root/reference/app/sdr/streamer.py:119-121 and the mlab.psd frame loop behind
// /root/reference/scripts/process_sigmf_data.py:188.  What changes is the data flow:
//
//   n = C*m + c,  m = 16a + b  (a, b in [0,16), c in [0,C))        X[k' + 256 k_c],  k' = k_a + 16 k_b
//   phase A : thread (b, c)   radix-16 over a   (input from the TMA-staged, 128B-swizzled frame; window fused)
//   phase B : thread (k_a, c) radix-16 over b   (twiddle W_256^{b k_a})       -> Y_c[k'] = FFT_256 of column c
//   phase C : thread j = k'   radix-C  over c   (twiddle W_N^{c k'})          -> fused epilogue (unchanged)
//
// The exchange A -> B only couples the 16 threads that share a column c.  A warp owns two whole columns
// (c = 2w, 2w+1), so that exchange is WARP-LOCAL: shared memory + __syncwarp, no CTA barrier, and the warps of a CTA
// drift apart freely through two thirds of a frame.  Only B -> C crosses warps: ONE barrier per frame (the exchange
// buffer is double-buffered by frame parity).  K1 needs two full barriers per frame with all 8 warps in lockstep.
//
// Everything is __host__ __device__: tests/emul re-executes these phases on the CPU (test infrastructure).
#pragma once
#include "spx_stft_device.cuh"

namespace spx {

template <int N>
struct Stft2Geom {
    static_assert(N == 1024 || N == 2048 || N == 4096, "K1v2 handles N = 256 * {4, 8, 16}");
    static constexpr int C = N / 256;          // columns = radix of the last pass
    static constexpr int T = N / 16;           // threads per frame (= 16 C)
    static constexpr int XS = 280;             // float2 per column in the exchange buffer: 16 x 17 padded tile, XS % 16 == 8
    static constexpr int X_F2 = C * XS;        // one exchange buffer
};

// 128-byte swizzle of the TMA tensor map (CU_TENSOR_MAP_SWIZZLE_128B): the 16-byte chunk index inside a 128-byte row is
// XORed with the row index mod 8
SPX_HD unsigned swz128(unsigned byte) { return byte ^ ((byte >> 3) & 0x70u); }

// phase-A thread (lane l of warp w) -> (b, c): a half-warp reads 8 rows x 2 columns of the swizzled frame, i.e. 16
// distinct 8-byte bank pairs (cf32: conflict-free; ci16: the two columns of a warp cover half of the banks, 2-way)
SPX_HD int k2_b_of(int tid) { return (((tid >> 4) & 1) << 3) | (tid & 7); }
SPX_HD int k2_c_of(int tid) { return ((tid >> 5) << 1) | ((tid >> 3) & 1); }

// window value slot: the table holds w[n] * scale for a < 8 laid out [a][tid]; by symmetry of np.hanning / np.blackman
// w[N-1-n] = w[n], and N-1-n maps (a, tid) -> (15-a, T-1-tid)
template <int N>
SPX_HD void k2_build_window(float* wtab, const float* win, int tid) {
    using G = Stft2Geom<N>;
    const int b = k2_b_of(tid), c = k2_c_of(tid);
#pragma unroll
    for (int a = 0; a < 8; ++a) wtab[a * G::T + tid] = ld_keep(win + 16 * G::C * a + G::C * b + c);
}

template <int TUNE>
SPX_HD void k2_dft16(float2* v) {
    if constexpr ((TUNE & TUNE_FMADFT) != 0) dft16_fma(v);
    else dft<16>(v);
}

// ---- phase A: staged input -> window -> radix-16 over a -> warp-local tile  X[c][17 b + k_a]
template <int N, int FMT, int TUNE>
SPX_HD void k2_phase_a(float2* v, int tid, const void* stage, const float* wtab, float2* X) {
    using G = Stft2Geom<N>;
    constexpr unsigned ELT = FMT == FMT_CF32 ? 8u : 4u;
    constexpr unsigned STEP = 16u * G::C * ELT;   // bytes between a and a+1
    const int b = k2_b_of(tid), c = k2_c_of(tid);
    const unsigned byte0 = ELT * (unsigned)(G::C * b + c);
    const char* st = reinterpret_cast<const char*>(stage);
#pragma unroll
    for (int a = 0; a < 16; ++a) {
        // STEP is a multiple of 1024 for N = 4096 (and 2048 cf32): the XOR term is then the same for every a and the
        // compiler folds a * STEP into the load's immediate offset
        const unsigned off = (STEP % 1024u == 0u) ? swz128(byte0) + (unsigned)a * STEP : swz128(byte0 + (unsigned)a * STEP);
        if (FMT == FMT_CF32) v[a] = *reinterpret_cast<const float2*>(st + off);
        else                 v[a] = ci16_to_f2<TUNE>(*reinterpret_cast<const unsigned int*>(st + off));
    }
    if (wtab != nullptr) {
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            const float w = a < 8 ? wtab[a * G::T + tid] : wtab[(15 - a) * G::T + (G::T - 1 - tid)];
            v[a].x *= w;
            v[a].y *= w;
        }
    }
    k2_dft16<TUNE>(v);
    float2* dst = X + G::XS * c + 17 * b;
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k] = v[k];
}

// ---- phase A with STRIP STAGING (N = 4096, hop = N / R, R = 2 or 4): the staging buffer is a ring of R hop-sized blocks; a
// frame inside a chunk only copies its newest block (1/R of the frame), the other R - 1 blocks are still there from its
// predecessors.  Logical block j of chunk-frame fi sits in ring slot (fi + j) mod R; `ring0` = fi mod R.
template <int FMT, int R, int TUNE>
SPX_HD void k2_phase_a_strip(float2* v, int tid, const void* stage, int ring0, const float* wtab, float2* X) {
    using G = Stft2Geom<4096>;
    constexpr unsigned ELT = FMT == FMT_CF32 ? 8u : 4u;
    constexpr unsigned STEP = 256u * ELT;            // bytes between registers a and a + 1 (a multiple of 1024)
    constexpr int APB = 16 / R;                      // registers per block
    constexpr unsigned BLK = APB * STEP;             // bytes per ring block (= hop samples)
    const int b = k2_b_of(tid), c = k2_c_of(tid);
    const char* st = reinterpret_cast<const char*>(stage) + swz128(ELT * (unsigned)(16 * b + c));
    unsigned boff[R];
#pragma unroll
    for (int j = 0; j < R; ++j) boff[j] = (unsigned)((j + ring0) & (R - 1)) * BLK;
#pragma unroll
    for (int a = 0; a < 16; ++a) {
        const unsigned off = boff[a / APB] + (unsigned)(a % APB) * STEP;
        if (FMT == FMT_CF32) v[a] = *reinterpret_cast<const float2*>(st + off);
        else                 v[a] = ci16_to_f2<TUNE>(*reinterpret_cast<const unsigned int*>(st + off));
    }
    if (wtab != nullptr) {
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            const float w = a < 8 ? wtab[a * G::T + tid] : wtab[(15 - a) * G::T + (G::T - 1 - tid)];
            v[a].x *= w;
            v[a].y *= w;
        }
    }
    k2_dft16<TUNE>(v);
    float2* dst = X + G::XS * c + 17 * b;
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k] = v[k];
}

// ---- phase B: radix-16 over b for (k_a, c); B1 = load + twiddle + DFT, B2 = store (a __syncwarp between them: the
// column tile is overwritten in place with Y_c[k_a + 16 k_b] at X[c][k_a + 16 k_b])
SPX_HD int k2_ka_of(int tid) { return tid & 15; }
SPX_HD int k2_cb_of(int tid) { return ((tid >> 5) << 1) | ((tid >> 4) & 1); }

template <int N, int TUNE = 0>
SPX_HD void k2_phase_b1(float2* v, int tid, const float2* X, const TwRegs<N>& twr) {
    using G = Stft2Geom<N>;
    const float2* src = X + G::XS * k2_cb_of(tid) + k2_ka_of(tid);
#pragma unroll
    for (int b = 0; b < 16; ++b) v[b] = src[17 * b];
    pass_twiddle_regs<N, 1>(v, twr);
    k2_dft16<TUNE>(v);
}
// same, with the 15 x 16 twiddles W_256^{b k_a} read from a shared-memory table tw256[(b - 1) * 16 + k_a] instead of being
// rebuilt from six register bases every frame: 15 conflict-free LDS.64 (both half-warps read the same 16 entries) against
// 9 complex products and 12 registers (used by K2v2, whose two roles leave the load/store pipe at 25 %)
template <int N, int TUNE = 0>
SPX_HD void k2_phase_b1_tab(float2* v, int tid, const float2* X, const float2* tw256) {
    using G = Stft2Geom<N>;
    const int ka = k2_ka_of(tid);
    const float2* src = X + G::XS * k2_cb_of(tid) + ka;
#pragma unroll
    for (int b = 0; b < 16; ++b) v[b] = src[17 * b];
#pragma unroll
    for (int b = 1; b < 16; ++b) v[b] = cmul(v[b], tw256[(b - 1) * 16 + ka]);
    k2_dft16<TUNE>(v);
}
template <int N>
SPX_HD void k2_phase_b2(const float2* v, int tid, float2* X) {
    using G = Stft2Geom<N>;
    float2* dst = X + G::XS * k2_cb_of(tid) + k2_ka_of(tid);
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[16 * k] = v[k];
}

// ---- phase C: radix-C over c for the butterflies j = tid + T u, then the fused epilogue of K1 (same register -> bin map)
template <int N, bool ACC, bool TW_IN_REGS, int TUNE = 0>
SPX_HD void k2_phase_c(float2* v, int tid, const float2* X, const StftParams& p, long long row, const float2* tw,
                       const TwRegs<N>& twr, StftAcc<ACC>& acc) {
    using G = Stft2Geom<N>;
    constexpr int R = G::C, NB = 16 / R;
#pragma unroll
    for (int u = 0; u < NB; ++u)
#pragma unroll
        for (int t = 0; t < R; ++t) v[u * R + t] = X[G::XS * t + tid + G::T * u];
    if constexpr (TW_IN_REGS) pass_twiddle_regs<N, 2>(v, twr);
    else pass_twiddle_table<N, 2, false>(v, tid, tw);
    if constexpr (R == 16 && (TUNE & TUNE_FMADFT) != 0) dft16_fma(v);
    else pass_dft<N, 2>(v);
    epilogue<N, ACC, TUNE>(v, tid, p, row, acc);
}

}  // namespace spx
