// N = 4096 instantiations of the fused STFT kernel (K1), including tuning variants.
#include "spx_stft2_kernel.cuh"

namespace spx {
int launch_stft_4k(StftLaunch& L) {
    switch (L.variant) {
        case 0:    // default = K1v2 with FMA-form DFTs and the FMA-pipe colormap index (variant 22); K1 when frames are not 128-byte aligned
        case 23:   // + strip staging of overlapping frames (experiment)
        case 22:   // + colormap index without the XU pipe
        case 21:   // K1v2 + FMA-form radix-16 DFTs
        case 20:   // K1v2: warp-local first exchange, one barrier per frame
            if (stft2_ok(L)) return launch_stft2(L);
            return launch_stft_n<4096, TW_REG, 2, true, TUNE_I2FP>(L);
        case 12:   // round-1 default (K1: two block barriers per frame, bulk-copy staging, register twiddle bases)
            return launch_stft_n<4096, TW_REG, 2, true, TUNE_I2FP>(L);  // default: TMA-staged input, register twiddle bases, int16 -> float on the ALU pipe
        case 8: return launch_stft_n<4096, TW_LDG, 2>(L);
        case 1: return launch_stft_n<4096, TW_LDG, 3>(L);
        case 2: return launch_stft_n<4096, TW_REG, 2>(L);
        case 3: return launch_stft_n<4096, TW_REG, 3>(L);
        case 4: return launch_stft_n<4096, TW_SMEM, 2>(L);
        case 5: return launch_stft_n<4096, TW_REG, 2, true>(L);   // same as 0
        case 6: return launch_stft_n<4096, TW_LDG, 2, true>(L);
        case 7: return launch_stft_n<4096, TW_HYB, 2, true>(L);   // pass-1 twiddles from smem, pass-2 from registers
        case 11: {  // software-pipelined rows-only kernel (falls back to the default when accumulators are wanted / unaligned)
            const bool acc = L.p.welch_acc != nullptr || L.p.maxhold != nullptr;
            if (!acc && stage_ok(L)) {
                return L.in_fmt == FMT_CF32 ? launch_stft_pipe_inst<4096, FMT_CF32, 2, TUNE_I2FP>(L)
                                            : launch_stft_pipe_inst<4096, FMT_CI16, 2, TUNE_I2FP>(L);
            }
            return launch_stft_n<4096, TW_REG, 2, true, TUNE_I2FP>(L);
        }
        case 10: return launch_stft_n<4096, TW_HYB, 2, true, TUNE_I2FP>(L);
        case 9: return launch_stft_n<4096, TW_REG, 2, true>(L);   // int16 -> float with I2F.S16 (XU pipe)
        default: return spx_set_error(SPX_E_INVALID, "unknown kernel variant %d for nfft 4096", L.variant);
    }
}
}  // namespace spx
