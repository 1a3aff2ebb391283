// N = 8192 instantiation of the fused STFT kernel (K1): 512 threads, 135 KiB of shared memory.
#include "spx_stft_kernel.cuh"

namespace spx {
int launch_stft_8k(StftLaunch& L) { return launch_stft_n<8192, TW_LDG, 1, true, TUNE_I2FP>(L); }
}  // namespace spx
