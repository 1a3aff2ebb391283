// spx_big2_device.cuh -- per-thread pieces of K2v2, the single-kernel 65536-point STFT (config 5 of BASELINE.json).
//
// N = 256 x 256 four-step (n = 256 n1 + n2, k = k1 + 256 k2).  Both steps are "16 column FFTs of 256 points over a
// [256 rows][16 columns] tile", i.e. exactly phases A and B of K1v2 (spx_stft2_device.cuh: swizzled TMA tile, radix-16
// over a, warp-local 16 x 16 exchange, radix-16 over b) -- no block barrier inside a transform:
//
//   role A (column tile g: n2 = 16 g + c): tile[n1][c] = x[256 n1 + n2] (TMA from the capture, window fused on load)
//        -> Y[k1] per column -> * W_N^{n2 k1} -> scratch T[c-major][k1]  (128-byte coalesced stores from registers)
//   role B (row tile g': k1 = 16 g' + r):   tile[n2][r] = T[n2][k1]       (TMA from the L2-resident scratch)
//        -> X[k1 + 256 k2] per row -> fused epilogue (|X|^2, Welch, max-hold, log2, uint8) -> 16-byte row runs
//
// A CTA alternates between the two roles (A of frame s, then B of frame s-1 of its lane of 16 CTAs), so the scratch of a
// frame is consumed one role after it was produced: 4 frames x lanes x 512 KiB stay resident in the 126 MB L2 and
// never travel to HBM.  Replaces np.fft.fft on 2^16-sample buffers (/root/reference/app/sdr/streamer.py:119,
// /root/reference/scripts/pyad-iio-test.py:50-61) plus the dB / waterfall lines (:121, callbacks.py:176-190).
#pragma once
#include "spx_stft2_device.cuh"

namespace spx {

enum { BIG2_N = 65536, BIG2_TILES = 16 };   // 16 column tiles / 16 row tiles of 16

// ---- phase A of either role, in two halves so that the block barrier that frees the staged tile can sit between them:
// (1) swizzled tile -> registers, window (role A), first butterfly layer of the radix-16 over a: every loaded value is
//     consumed by an arithmetic instruction here, so all shared-memory reads of the tile are complete when this returns;
// (2) second layer -> warp-local tile X[c][17 b + k_a].
// `w` (role A): this thread's 16 window * scale values, w[WS * a] (a table [16 a][256 tid] offset by tid: WS = 256; or a
// private array: WS = 1); nullptr for role B / rect window.
template <int WS = 256>
SPX_HD void big2_load_tile(float2* v, int tid, const void* stage, const float* w) {
    const int b = k2_b_of(tid), c = k2_c_of(tid);
    const unsigned off0 = swz128(8u * (unsigned)(16 * b + c));
    const char* st = reinterpret_cast<const char*>(stage);
#pragma unroll
    for (int a = 0; a < 16; ++a) v[a] = *reinterpret_cast<const float2*>(st + off0 + (unsigned)a * 2048u);
    if (w != nullptr) {
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            const float wa = w[WS * a];
            v[a].x *= wa;
            v[a].y *= wa;
        }
    }
    dft16_layer1(v);
}
// The same for int16 input (role A only): the TMA box is {64 B x 256 rows} with CU_TENSOR_MAP_SWIZZLE_64B (the 16-byte chunk
// index inside a 64-byte row is XORed with bits 7..8 of the address, i.e. with (row >> 1) & 3): the 8 rows a half-warp
// touches fall on 8 distinct (128-byte half, chunk) positions, 2-way conflicts between its two halves as in K1v2.
SPX_HD unsigned swz64(unsigned byte) { return byte ^ ((byte >> 3) & 0x30u); }
template <int WS = 256>
SPX_HD void big2_load_tile_ci16(float2* v, int tid, const void* stage, const float* w) {
    const int b = k2_b_of(tid), c = k2_c_of(tid);
    const unsigned off0 = swz64(4u * (unsigned)(16 * b + c));   // 64 (16 a + b) + 4 c: a only moves bits >= 10
    const char* st = reinterpret_cast<const char*>(stage);
#pragma unroll
    for (int a = 0; a < 16; ++a) v[a] = ci16_to_f2<TUNE_I2FP>(*reinterpret_cast<const unsigned int*>(st + off0 + (unsigned)a * 1024u));
    if (w != nullptr) {
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            const float wa = w[WS * a];
            v[a].x *= wa;
            v[a].y *= wa;
        }
    }
    dft16_layer1(v);
}
SPX_HD void big2_dft_store(float2* v, int tid, float2* X) {
    using G = Stft2Geom<4096>;
    dft16_fma_layer2(v);
    float2* dst = X + G::XS * k2_c_of(tid) + 17 * k2_b_of(tid);
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k] = v[k];
}
template <int TUNE, int WS = 256>
SPX_HD void big2_phase_a(float2* v, int tid, const void* stage, const float* w, float2* X) {
    static_assert((TUNE & TUNE_FMADFT) != 0, "K2v2 uses the FMA-form radix-16");
    big2_load_tile<WS>(v, tid, stage, w);
    big2_dft_store(v, tid, X);
}

// sample index of register a of phase-A thread tid in column tile g (host: builds the [16 g][16 a][256 tid] window table)
SPX_HD int big2_sample_of(int g, int a, int tid) { return 256 * (16 * a + k2_b_of(tid)) + 16 * g + k2_c_of(tid); }

// ---- role A output: Y[k1 = k_a + 16 k_b] of column c (phase-B thread map: lane = 16 (c & 1) + k_a) times
// W_N^{n2 k1}.  The 16 twiddles of a thread are products of 7 frame-invariant bases kept in shared memory
// (`bases` = this tile's [7][256] float2 table, built on the host in float64):  beta_t = W^{n2 (k_a + 16 t)}, t = 0..3;  gamma_q = W^{64 n2 q}, q = 1..3.
template <int TUNE>
SPX_HD void big2_twiddle_store(float2* v, int tid, const float2* bases, float2* t_tile /* this CTA's [16 c][256 k1] block */) {
    const float2 b0 = bases[0 * 256 + tid], b1 = bases[1 * 256 + tid], b2 = bases[2 * 256 + tid], b3 = bases[3 * 256 + tid];
    const float2 g1 = bases[4 * 256 + tid], g2 = bases[5 * 256 + tid], g3 = bases[6 * 256 + tid];
    v[0] = cmul(v[0], b0);   v[1] = cmul(v[1], b1);   v[2] = cmul(v[2], b2);   v[3] = cmul(v[3], b3);
    v[4] = cmul(v[4], cmul(g1, b0));   v[5] = cmul(v[5], cmul(g1, b1));   v[6] = cmul(v[6], cmul(g1, b2));   v[7] = cmul(v[7], cmul(g1, b3));
    v[8] = cmul(v[8], cmul(g2, b0));   v[9] = cmul(v[9], cmul(g2, b1));   v[10] = cmul(v[10], cmul(g2, b2)); v[11] = cmul(v[11], cmul(g2, b3));
    v[12] = cmul(v[12], cmul(g3, b0)); v[13] = cmul(v[13], cmul(g3, b1)); v[14] = cmul(v[14], cmul(g3, b2)); v[15] = cmul(v[15], cmul(g3, b3));
    float2* dst = t_tile + 256 * k2_cb_of(tid) + k2_ka_of(tid);   // a half-warp writes 16 consecutive k1 = 128 bytes
#pragma unroll
    for (int kb = 0; kb < 16; ++kb) dst[16 * kb] = v[kb];
}

// host-side table of the bases above for every (tile g, phase-B thread tid): [16 g][7][256]
SPX_HD void big2_base_exponents(int g, int tid, unsigned* e /*[7]*/) {
    const unsigned n2 = (unsigned)(16 * g + k2_cb_of(tid)), ka = (unsigned)k2_ka_of(tid);
    for (unsigned t = 0; t < 4; ++t) e[t] = (n2 * (ka + 16u * t)) & 65535u;
    for (unsigned q = 1; q < 4; ++q) e[3 + q] = (64u * n2 * q) & 65535u;
}

// ---- role B output: thread (k_a, r) (same lane map, r = row inside the tile) holds X[k1 + 256 k2], k1 = 16 g + r,
// k2 = k_a + 16 k_b.  |X|^2, accumulate, log2 (eps-exact form when needed), uint8 index into the [256 k2s][16 r] byte tile
// (k2s = (k2 + 128) mod 256 is the fftshift position of the run).
template <bool ACC, int TUNE>
SPX_HD void big2_epilogue(float2* v, int tid, float db_eps, float db_pw_min, float q_a, float q_b, bool want_rows,
                          StftAcc<ACC>& acc, unsigned char* u8tile) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i].x = v[i].x * v[i].x + v[i].y * v[i].y;
    if (ACC) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            acc.sum[i] += v[i].x;
            acc.mx[i] = fmaxf(acc.mx[i], v[i].x);
        }
    }
    if (!want_rows) return;
    float pmin = v[0].x;
#pragma unroll
    for (int i = 1; i < 16; ++i) pmin = fminf(pmin, v[i].x);
    if (pmin >= db_pw_min) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i].y = fast_log2(v[i].x);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i].y = 2.0f * fast_log2(fast_sqrt(v[i].x) + db_eps);
    }
    const int ka = k2_ka_of(tid), r = k2_cb_of(tid);
    const float qa256 = q_a * 0.00390625f, qb256 = q_b * 0.00390625f;
#pragma unroll
    for (int kb = 0; kb < 16; ++kb) {
        const int k2s = (ka + 16 * kb + 128) & 255;
        unsigned int q;
        if constexpr ((TUNE & TUNE_QFMA) != 0) q = sat_floor_u8_fma(v[kb].y, qa256, qb256);
        else q = sat_floor_u8(quant_pre(v[kb].y, q_a, q_b));
        u8tile[16 * k2s + r] = (unsigned char)q;
    }
}

// fftshift position of accumulator register kb of thread tid in row tile g
SPX_HD int big2_acc_pos(int g, int tid, int kb) {
    const int k2s = (k2_ka_of(tid) + 16 * kb + 128) & 255;
    return 16 * g + k2_cb_of(tid) + 256 * k2s;
}

}  // namespace spx
