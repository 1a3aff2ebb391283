// N = 1024 / 2048 / 4096 instantiations of K1v2 (warp-local first exchange, one block barrier per frame, swizzled TMA
// tensor staging).  See spx_stft2_kernel.cuh / spx_stft2_device.cuh.
#include "spx_stft2_kernel.cuh"

namespace spx {

tmap_encode_fn tmap_encoder() {
    static tmap_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (tmap_encode_fn)p;
        else
            cudaGetLastError();
        tried = true;
    }
    return fn;
}

int launch_stft2(StftLaunch& L) {
    if (L.variant == 23) {   // strip staging of overlapping frames (experiment; N = 4096, hop = N/2 or N/4), else the default
        int rc = SPX_OK;
        if (launch_stft2_strip<2, TUNE_I2FP | TUNE_FMADFT | TUNE_QFMA>(L, &rc)) return rc;
        return launch_stft2_n<4096, 2, TUNE_I2FP | TUNE_FMADFT | TUNE_QFMA>(L);
    }
    if (L.variant == 0 && L.nfft == 4096 && L.in_fmt == FMT_CF32 && L.p.hop == 1024) {
        // measured (profiles/r02_strip_staging_sweep.jsonl): strip staging wins 3 % on cf32 at 75 % overlap (the frame is
        // 32 KB, three quarters of it already in shared memory), is neutral for ci16 and at 50 % overlap -- used only here
        int rc = SPX_OK;
        if (launch_stft2_strip<2, TUNE_I2FP | TUNE_FMADFT | TUNE_QFMA>(L, &rc)) return rc;
    }
    if (L.variant == 0 && L.nfft == 4096 && L.in_fmt == FMT_CF32 && L.p.hop >= 4096) {
        // measured (profiles/r02_l2_prefetch_sweep.txt): the L2 prefetch of the next frame is +5 % here (one 32 KB frame in flight
        // per CTA, all of it from HBM), neutral for int16, -1 % with overlap, -3 % for N = 2048 / 1024 (2 / 4 frames in flight per CTA)
        const bool acc = L.p.welch_acc != nullptr || L.p.maxhold != nullptr;
        constexpr int T = TUNE_I2FP | TUNE_FMADFT | TUNE_QFMA | TUNE_L2PF;
        return acc ? launch_stft2_inst<4096, FMT_CF32, true, 2, T>(L) : launch_stft2_inst<4096, FMT_CF32, false, 2, T>(L);
    }
    if (L.variant == 22 || L.variant == 0) {   // default: FMA-form DFTs + uint8 index on the FMA / ALU pipes instead of F2I (XU)
        switch (L.nfft) {
            case 1024: return launch_stft2_n<1024, 2, TUNE_I2FP | TUNE_FMADFT | TUNE_QFMA>(L);
            case 2048: return launch_stft2_n<2048, 2, TUNE_I2FP | TUNE_FMADFT | TUNE_QFMA>(L);
            case 4096: return launch_stft2_n<4096, 2, TUNE_I2FP | TUNE_FMADFT | TUNE_QFMA>(L);
        }
    }
    if (L.variant == 21) {   // radix-16 DFTs in fused-multiply-add form (dft16_fma)
        switch (L.nfft) {
            case 1024: return launch_stft2_n<1024, 2, TUNE_I2FP | TUNE_FMADFT>(L);
            case 2048: return launch_stft2_n<2048, 2, TUNE_I2FP | TUNE_FMADFT>(L);
            case 4096: return launch_stft2_n<4096, 2, TUNE_I2FP | TUNE_FMADFT>(L);
        }
    }
    switch (L.nfft) {
        case 1024: return launch_stft2_n<1024, 2, TUNE_I2FP>(L);
        case 2048: return launch_stft2_n<2048, 2, TUNE_I2FP>(L);
        case 4096: return launch_stft2_n<4096, 2, TUNE_I2FP>(L);
        default: return spx_set_error(SPX_E_UNSUPPORTED, "K1v2 handles nfft 1024 / 2048 / 4096, not %d", L.nfft);
    }
}

}  // namespace spx
