// spx_fft_core.cuh -- register-level Stockham FFT building blocks (sm_100a).
//
// Replaces the `np.fft.fft(samples)` call of the reference hot path
// (/root/reference/app/sdr/streamer.py:119): forward DFT, exp(-2*pi*i*k*n/N), unnormalised.
//
// Design (see DESIGN.md "K1"):
//   * N = R0*R1*...  with R0 = 16 and every radix in {2,4,8,16}; every thread owns 16 points
//     per pass, T = N/16 threads cooperate on one frame.
//   * Stockham autosort: pass s reads in[j + t*N/R] (lane-contiguous, conflict-free), applies
//     the DIT twiddle W_{Ns*R}^{(j mod Ns) t}, does the radix-R DFT in registers and writes
//     out[(j/Ns)*Ns*R + (j mod Ns) + t*Ns].  The last pass leaves natural-order bins in
//     registers for the fused epilogue -- no bit reversal, no extra pass.
//   * Exchanges go through shared memory as float2 (64-bit accesses are served per half-warp,
//     so Ns >= 16 makes every write conflict-free); only the first exchange (Ns = 1) needs the
//     17/16 padding `pad0`.
//
// All functions are __host__ __device__ so that tests/emul (a CPU re-execution of the same
// per-thread code, test infrastructure only) can check indices and accuracy without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef SPX_HD
#define SPX_HD __host__ __device__ __forceinline__
#endif

namespace spx {

// ------------------------------------------------------------------ complex helpers
// complex add / subtract: one packed FADD2 on sm_100a (same rounding as two scalar adds)
#ifndef SPX_PACKED_F32X2
#define SPX_PACKED_F32X2 1
#endif
SPX_HD float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && SPX_PACKED_F32X2
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
SPX_HD float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && SPX_PACKED_F32X2
    return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
SPX_HD float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * (-i)  and a * (+i)
SPX_HD float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }
SPX_HD float2 mul_pi(float2 a) { return make_float2(-a.y, a.x); }

// W_16^k = cos(2 pi k/16) - i sin(2 pi k/16), k = 0..15 (float-rounded from double)
#define SPX_C1 0.92387953251128673848f
#define SPX_S1 0.38268343236508978178f
#define SPX_H  0.70710678118654752440f

template <int K16>
SPX_HD float2 mul_w16(float2 a) {
    constexpr int k = ((K16 % 16) + 16) % 16;
    if constexpr (k == 0) return a;
    else if constexpr (k == 4) return mul_mi(a);
    else if constexpr (k == 8) return make_float2(-a.x, -a.y);
    else if constexpr (k == 12) return mul_pi(a);
    else if constexpr (k == 2) return make_float2((a.x + a.y) * SPX_H, (a.y - a.x) * SPX_H);
    else if constexpr (k == 6) return make_float2((a.y - a.x) * SPX_H, -(a.x + a.y) * SPX_H);
    else if constexpr (k == 10) return make_float2(-(a.x + a.y) * SPX_H, (a.x - a.y) * SPX_H);
    else if constexpr (k == 14) return make_float2((a.x - a.y) * SPX_H, (a.x + a.y) * SPX_H);
    else {
        // generic: (c - i s) with c = cos(2 pi k/16), s = sin(2 pi k/16)
        constexpr float c = (k == 1 || k == 15) ? SPX_C1 : (k == 3 || k == 13) ? SPX_S1
                          : (k == 5 || k == 11) ? -SPX_S1 : /* 7, 9 */ -SPX_C1;
        constexpr float s = (k == 1 || k == 7) ? SPX_S1 : (k == 3 || k == 5) ? SPX_C1
                          : (k == 9 || k == 15) ? -SPX_S1 : /* 11, 13 */ -SPX_C1;
        return make_float2(a.x * c + a.y * s, a.y * c - a.x * s);
    }
}

// ------------------------------------------------------------------ in-register DFTs (natural order out)
SPX_HD void dft2(float2& a, float2& b) {
    float2 t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

SPX_HD void dft4(float2& v0, float2& v1, float2& v2, float2& v3) {
    const float2 t0 = cadd(v0, v2), t1 = csub(v0, v2);
    const float2 t2 = cadd(v1, v3);
    const float2 u3 = make_float2(v1.y - v3.y, v3.x - v1.x);  // -i (v1 - v3), built directly in rotated form
    v0 = cadd(t0, t2);
    v2 = csub(t0, t2);
    v1 = cadd(t1, u3);  // t1 - i t3
    v3 = csub(t1, u3);  // t1 + i t3
}

template <int R>
SPX_HD void dft(float2* v);

template <>
SPX_HD void dft<2>(float2* v) { dft2(v[0], v[1]); }

template <>
SPX_HD void dft<4>(float2* v) { dft4(v[0], v[1], v[2], v[3]); }

template <>
SPX_HD void dft<8>(float2* v) {
    // 8 = 2 x 4 : Y[c][b] = x[b] +- x[b+4]; Y[1][b] *= W_8^b; DFT4 over b -> X[c + 2d]
    dft2(v[0], v[4]);
    dft2(v[1], v[5]);
    dft2(v[2], v[6]);
    dft2(v[3], v[7]);
    v[5] = mul_w16<2>(v[5]);
    v[6] = mul_w16<4>(v[6]);
    v[7] = mul_w16<6>(v[7]);
    dft4(v[0], v[1], v[2], v[3]);  // -> X[0], X[2], X[4], X[6]
    dft4(v[4], v[5], v[6], v[7]);  // -> X[1], X[3], X[5], X[7]
    float2 x1 = v[4], x3 = v[5], x5 = v[6], x7 = v[7];
    float2 x2 = v[1], x4 = v[2], x6 = v[3];
    v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
}

template <>
SPX_HD void dft<16>(float2* v) {
    // 16 = 4 x 4 : inner DFT4 over a' (stride 4) for each b -> Y[c'][b] (stored at v[4c'+b]),
    // twiddle W_16^{b c'}, outer DFT4 over b -> X[c' + 4d]
    dft4(v[0], v[4], v[8], v[12]);
    dft4(v[1], v[5], v[9], v[13]);
    dft4(v[2], v[6], v[10], v[14]);
    dft4(v[3], v[7], v[11], v[15]);
    v[5] = mul_w16<1>(v[5]);   v[6] = mul_w16<2>(v[6]);    v[7] = mul_w16<3>(v[7]);
    v[9] = mul_w16<2>(v[9]);   v[10] = mul_w16<4>(v[10]);  v[11] = mul_w16<6>(v[11]);
    v[13] = mul_w16<3>(v[13]); v[14] = mul_w16<6>(v[14]);  v[15] = mul_w16<9>(v[15]);
    dft4(v[0], v[1], v[2], v[3]);      // c'=0 -> X[0], X[4], X[8],  X[12]
    dft4(v[4], v[5], v[6], v[7]);      // c'=1 -> X[1], X[5], X[9],  X[13]
    dft4(v[8], v[9], v[10], v[11]);    // c'=2 -> X[2], X[6], X[10], X[14]
    dft4(v[12], v[13], v[14], v[15]);  // c'=3 -> X[3], X[7], X[11], X[15]
    // v[4c'+d] holds X[c'+4d]  ->  transpose the 4x4 register tile (pure renaming)
    float2 t;
#define SPX_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
    SPX_SWAP(1, 4) SPX_SWAP(2, 8) SPX_SWAP(3, 12) SPX_SWAP(6, 9) SPX_SWAP(7, 13) SPX_SWAP(11, 14)
#undef SPX_SWAP
}

// ------------------------------------------------------------------ radix-16 DFT in fused-multiply-add form
// Same 4 x 4 decomposition as dft<16>, but no W_16 twiddle is ever applied as a stand-alone multiply (Linzer-Feig
// style): a factor c (1 - i t) is split into the two-FMA rotation (x + t y, y - t x) and the real scale c, and the
// scale (cos(pi/8) or 1/sqrt 2) rides on the next butterfly as the multiplier of an FMA:  a +- c * b'.
// 144 FP32 operations instead of 160 (64 + 16 + 22 + 20 + 22), the outer butterflies are packed FFMA2 on sm_100a.
// Natural-order output, like dft<16>.
SPX_HD float2 cfma(float2 b, float s, float2 a) {   // a + s * b
#if defined(__CUDA_ARCH__) && SPX_PACKED_F32X2
    return __ffma2_rn(b, make_float2(s, s), a);
#else
    return make_float2(fmaf(b.x, s, a.x), fmaf(b.y, s, a.y));
#endif
}
#define SPX_T1 0.41421356237309504880f   // tan(pi/8)

// first layer (four radix-4 butterflies over stride 4): consumes every input once -- K2v2 runs it in front of its block
// barrier so that all shared-memory reads of the staged tile are complete there
SPX_HD void dft16_layer1(float2* v) {
    dft4(v[0], v[4], v[8], v[12]);
    dft4(v[1], v[5], v[9], v[13]);
    dft4(v[2], v[6], v[10], v[14]);
    dft4(v[3], v[7], v[11], v[15]);
}
SPX_HD void dft16_fma_layer2(float2* v);
SPX_HD void dft16_fma(float2* v) {
    dft16_layer1(v);
    dft16_fma_layer2(v);
}
SPX_HD void dft16_fma_layer2(float2* v) {
    // c' = 0: no twiddles
    dft4(v[0], v[1], v[2], v[3]);
    {   // c' = 1: y1 = v5 W^1 = C1 (x + t y, y - t x);  y2 = v6 W^2 = H (x + y, y - x);  y3 = v7 W^3 = C1 * (-i) (x - t y, y + t x)
        const float2 p1 = make_float2(fmaf(SPX_T1, v[5].y, v[5].x), fmaf(-SPX_T1, v[5].x, v[5].y));
        const float2 q3 = make_float2(fmaf(-SPX_T1, v[7].y, v[7].x), fmaf(SPX_T1, v[7].x, v[7].y));
        const float2 p2 = make_float2(v[6].x + v[6].y, v[6].y - v[6].x);
        const float2 t0 = cfma(p2, SPX_H, v[4]), t1 = cfma(p2, -SPX_H, v[4]);
        // y3 / C1 = (q3.y, -q3.x);  t2' = p1 + y3/C1;  t3' = p1 - y3/C1;  u3' = -i t3' built directly in rotated form
        const float2 t2 = make_float2(p1.x + q3.y, p1.y - q3.x);
        const float2 u3 = make_float2(p1.y + q3.x, q3.y - p1.x);
        v[4] = cfma(t2, SPX_C1, t0);
        v[6] = cfma(t2, -SPX_C1, t0);
        v[5] = cfma(u3, SPX_C1, t1);
        v[7] = cfma(u3, -SPX_C1, t1);
    }
    {   // c' = 2: y1 = v9 W^2 = H (x + y, y - x);  y2 = v10 W^4 = (y, -x);  y3 = v11 W^6 = H (y - x, -(x + y))
        const float2 p1 = make_float2(v[9].x + v[9].y, v[9].y - v[9].x);
        const float2 p3 = make_float2(v[11].y - v[11].x, -(v[11].x + v[11].y));
        const float2 t0 = make_float2(v[8].x + v[10].y, v[8].y - v[10].x);
        const float2 t1 = make_float2(v[8].x - v[10].y, v[8].y + v[10].x);
        const float2 t2 = cadd(p1, p3);
        const float2 u3 = make_float2(p1.y - p3.y, p3.x - p1.x);      // -i (p1 - p3), built directly in rotated form
        v[8] = cfma(t2, SPX_H, t0);
        v[10] = cfma(t2, -SPX_H, t0);
        v[9] = cfma(u3, SPX_H, t1);
        v[11] = cfma(u3, -SPX_H, t1);
    }
    {   // c' = 3: y1 = v13 W^3 = C1 * (-i) (x - t y, y + t x);  y2 = v14 W^6 = H (y - x, -(x + y));  y3 = v15 W^9 = -C1 (x + t y, y - t x)
        const float2 q1 = make_float2(fmaf(-SPX_T1, v[13].y, v[13].x), fmaf(SPX_T1, v[13].x, v[13].y));
        const float2 q3 = make_float2(fmaf(SPX_T1, v[15].y, v[15].x), fmaf(-SPX_T1, v[15].x, v[15].y));
        const float2 p2 = make_float2(v[14].y - v[14].x, -(v[14].x + v[14].y));
        const float2 t0 = cfma(p2, SPX_H, v[12]), t1 = cfma(p2, -SPX_H, v[12]);
        // y1 / C1 = (q1.y, -q1.x);  y3 / C1 = -q3
        const float2 t2 = make_float2(q1.y - q3.x, -q1.x - q3.y);      // (y1 + y3) / C1
        const float2 u3 = make_float2(q3.y - q1.x, -(q1.y + q3.x));    // -i (y1 - y3) / C1, directly in rotated form
        v[12] = cfma(t2, SPX_C1, t0);
        v[14] = cfma(t2, -SPX_C1, t0);
        v[13] = cfma(u3, SPX_C1, t1);
        v[15] = cfma(u3, -SPX_C1, t1);
    }
    // v[4c'+d] holds X[c'+4d]  ->  transpose the 4x4 register tile (pure renaming)
    float2 t;
#define SPX_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
    SPX_SWAP(1, 4) SPX_SWAP(2, 8) SPX_SWAP(3, 12) SPX_SWAP(6, 9) SPX_SWAP(7, 13) SPX_SWAP(11, 14)
#undef SPX_SWAP
}

// ------------------------------------------------------------------ compile-time plan
// Radix of pass s for an N-point transform (0 when s >= number of passes).
SPX_HD constexpr int plan_radix(int n, int s) {
    // first radix is always 16; the remaining factor N/16 is split greedily into 16s then the rest,
    // except 8192 = 16*8*8*8 and 512 = 16*8*4 style splits which keep radices balanced.
    int rest = n / 16;
    if (s == 0) return n >= 16 ? 16 : 0;
    int r[4] = {0, 0, 0, 0};
    int cnt = 0;
    // factor `rest` into radices <= 16, largest first but avoid a trailing radix-2 when possible
    while (rest > 1) {
        int f = rest >= 16 ? 16 : rest;
        if (rest == 32) f = 8;   // 32 = 8*4 instead of 16*2
        if (rest == 512) f = 8;  // 512 = 8*8*8 instead of 16*16*2
        r[cnt++] = f;
        rest /= f;
    }
    return (s - 1) < cnt ? r[s - 1] : 0;
}
SPX_HD constexpr int plan_passes(int n) {
    int p = 0;
    while (p < 5 && plan_radix(n, p) != 0) ++p;
    return p;
}
// product of radices of passes < s
SPX_HD constexpr int plan_ns(int n, int s) {
    int ns = 1;
    for (int q = 0; q < s; ++q) ns *= plan_radix(n, q);
    return ns;
}
// offset (in float2) of pass s (s >= 1) inside the twiddle table; layout [t-1][jm], jm in [0,Ns)
SPX_HD constexpr int plan_tw_offset(int n, int s) {
    int off = 0;
    for (int q = 1; q < s; ++q) off += plan_ns(n, q) * (plan_radix(n, q) - 1);
    return off;
}
SPX_HD constexpr int plan_tw_size(int n) { return plan_tw_offset(n, plan_passes(n)); }

// padding of the first exchange buffer (after the Ns = 1 pass): one float2 per 16
SPX_HD constexpr int pad0(int i) { return i + (i >> 4); }
SPX_HD constexpr int padded_size(int n) { return n + (n >> 4); }

}  // namespace spx
