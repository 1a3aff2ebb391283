// N = 1024, 2048 instantiations of the fused STFT kernel (K1).
#include "spx_stft_kernel.cuh"

namespace spx {
int launch_stft_1k2k(StftLaunch& L) {
    switch (L.nfft) {
        case 1024: return launch_stft_n<1024, TW_LDG, 2, true, TUNE_I2FP>(L);
        case 2048: return launch_stft_n<2048, TW_LDG, 2, true, TUNE_I2FP>(L);
        default: return spx_set_error(SPX_E_UNSUPPORTED, "nfft %d", L.nfft);
    }
}
}  // namespace spx
