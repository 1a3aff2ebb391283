// N = 16 .. 512 instantiations of the fused STFT kernel (K1).
#include "spx_stft_kernel.cuh"

namespace spx {
int launch_stft_small(StftLaunch& L) {
    switch (L.nfft) {
        case 16: return launch_stft_n<16, TW_LDG, 2, false, TUNE_I2FP>(L);
        case 32: return launch_stft_n<32, TW_LDG, 2, false, TUNE_I2FP>(L);
        case 64: return launch_stft_n<64, TW_LDG, 2, false, TUNE_I2FP>(L);
        case 128: return launch_stft_n<128, TW_LDG, 2, false, TUNE_I2FP>(L);
        case 256: return launch_stft_n<256, TW_LDG, 2, false, TUNE_I2FP>(L);
        case 512: return launch_stft_n<512, TW_LDG, 2, true, TUNE_I2FP>(L);
        default: return spx_set_error(SPX_E_UNSUPPORTED, "nfft %d", L.nfft);
    }
}
}  // namespace spx
