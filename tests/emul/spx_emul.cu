// spx_emul.cu -- TEST INFRASTRUCTURE ONLY.  Re-executes the per-thread phases of the CUDA STFT kernel
// (sdr_iq_visualizer_b200/csrc/spx_stft_device.cuh, the exact code the GPU runs) on the CPU, thread
// by thread and barrier by barrier, so that the `-m "not gpu"` tests can check the Stockham index
// maps, padding, twiddle tables and epilogue against the numpy oracle without a GPU.
// It is never linked into libspx and never imported by the product package.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../sdr_iq_visualizer_b200/csrc/spx_stft_device.cuh"
#include "../../sdr_iq_visualizer_b200/csrc/spx_tables.h"

using namespace spx;

static bool g_staged = false;  // emulate the bulk-copy staging buffer (pass 0 reads shared memory)

template <int N, int FMT, bool ACC, int TWM, int S>
static void run_phase(std::vector<float2>& v, StftParams& p, long long s0, long long row, float2* A, float2* B,
                      const float2* tw, std::vector<TwRegs<N>>& twr, std::vector<StftAcc<ACC>>& acc) {
    constexpr int T = N / 16;
    std::vector<unsigned char> stage;
    const void* st = nullptr;
    if (S == 0 && g_staged) {
        const size_t elt = FMT == FMT_CF32 ? 8 : 4;
        stage.resize((size_t)N * elt);
        memcpy(stage.data(), (const char*)p.in + (size_t)s0 * elt, stage.size());
        st = stage.data();
    }
    for (int tid = 0; tid < T; ++tid)
        stft_phase<N, FMT, ACC, TWM, S>(&v[(size_t)tid * 16], tid, p, s0, row, true, A, B, tw, twr[tid], acc[tid], st,
                                        st ? p.win : nullptr);  // first N/2 entries of the global table = the half table
}

template <int N, int FMT, bool ACC, int TWM>
static int emul_stft(StftParams p) {
    constexpr int T = N / 16, P = plan_passes(N);
    std::vector<float2> tw = build_twiddles(N);
    p.tw = tw.data();
    std::vector<float2> A((size_t)padded_size(N)), B((size_t)N), v((size_t)T * 16);
    std::vector<TwRegs<N>> twr((size_t)T);
    std::vector<StftAcc<ACC>> acc((size_t)T);
    if constexpr (TWM == TW_REG && P > 1) {
        for (int tid = 0; tid < T; ++tid) {
            if constexpr (P > 1) tw_regs_load_pass<N, 1>(twr[tid], tid, p.tw);
            if constexpr (P > 2) tw_regs_load_pass<N, 2>(twr[tid], tid, p.tw);
            if constexpr (P > 3) tw_regs_load_pass<N, 3>(twr[tid], tid, p.tw);
        }
    }
    for (auto& a : acc) a.reset();
    const long long F = p.frames_per_stream;
    for (long long chunk = 0; chunk < p.total_chunks; ++chunk) {
        const long long stream = chunk / p.chunks_per_stream;
        const long long f0 = (chunk - stream * p.chunks_per_stream) * p.frames_per_chunk;
        long long nf = F - f0;
        if (nf > p.frames_per_chunk) nf = p.frames_per_chunk;
        for (long long fi = 0; fi < nf; ++fi) {
            const long long s0 = stream * p.stream_stride + (f0 + fi) * p.hop;
            const long long row = stream * F + f0 + fi;
            run_phase<N, FMT, ACC, TWM, 0>(v, p, s0, row, A.data(), B.data(), p.tw, twr, acc);
            if constexpr (P > 1) run_phase<N, FMT, ACC, TWM, 1>(v, p, s0, row, A.data(), B.data(), p.tw, twr, acc);
            if constexpr (P > 2) run_phase<N, FMT, ACC, TWM, 2>(v, p, s0, row, A.data(), B.data(), p.tw, twr, acc);
            if constexpr (P > 3) run_phase<N, FMT, ACC, TWM, 3>(v, p, s0, row, A.data(), B.data(), p.tw, twr, acc);
        }
        if constexpr (ACC) {
            for (int tid = 0; tid < T; ++tid) {
                for (int i = 0; i < 16; ++i) {
                    const long long o = stream * N + acc_pos<N>(tid, i);
                    if (p.welch_acc) p.welch_acc[o] += (double)acc[tid].sum[i];
                    if (p.maxhold && acc[tid].mx[i] > p.maxhold[o]) p.maxhold[o] = acc[tid].mx[i];
                }
                acc[tid].reset();
            }
        }
    }
    return 0;
}

template <int N, int TWM>
static int emul_n(int fmt, bool acc, StftParams& p) {
    if (fmt == FMT_CF32) return acc ? emul_stft<N, FMT_CF32, true, TWM>(p) : emul_stft<N, FMT_CF32, false, TWM>(p);
    return acc ? emul_stft<N, FMT_CI16, true, TWM>(p) : emul_stft<N, FMT_CI16, false, TWM>(p);
}

extern "C" int spx_emul_stft(int nfft, int in_fmt, int tw_mode, const void* in, long long n_samples, int n_streams,
                             long long stream_stride, int hop, const float* win, float db_eps, float vmin, float vmax,
                             int frames_per_chunk, float* db_rows, unsigned char* wf_rows, float* spec_rows,
                             double* welch_acc, float* maxhold) {
    StftParams p;
    memset(&p, 0, sizeof(p));
    p.in = in;
    p.stream_stride = stream_stride;
    p.n_streams = n_streams;
    p.hop = hop;
    p.frames_per_stream = n_samples < nfft ? 0 : (n_samples - nfft) / hop + 1;
    p.win = win;
    p.db_rows = db_rows;
    p.wf_rows = wf_rows;
    p.spec_rows = reinterpret_cast<float2*>(spec_rows);
    p.welch_acc = welch_acc;
    p.maxhold = maxhold;
    p.db_eps = db_eps;
    p.db_pw_min = db_eps * db_eps * 1099511627776.0f;  // (2^20 eps)^2
    p.q_a = (float)(3.01029995663981195214 * 256.0 / ((double)vmax - (double)vmin));  // 10 log10(2) * scale
    p.q_b = (float)(-(double)vmin * 256.0 / ((double)vmax - (double)vmin));
    if (p.frames_per_stream == 0) return 0;
    p.frames_per_chunk = frames_per_chunk;
    p.chunks_per_stream = (int)((p.frames_per_stream + frames_per_chunk - 1) / frames_per_chunk);
    p.total_chunks = (long long)p.chunks_per_stream * n_streams;
    const bool acc = welch_acc != nullptr || maxhold != nullptr;
    g_staged = (tw_mode & 0x10) != 0;
    tw_mode &= 0xf;
#define CASE(NN)                                                                  \
    case NN:                                                                      \
        if (tw_mode == TW_REG) {                                                  \
            if constexpr (NN == 256 || NN == 4096) return emul_n<NN, TW_REG>(in_fmt, acc, p); \
            return -4;                                                            \
        }                                                                         \
        return emul_n<NN, TW_LDG>(in_fmt, acc, p);
    switch (nfft) {
        CASE(16) CASE(32) CASE(64) CASE(128) CASE(256) CASE(512) CASE(1024) CASE(2048) CASE(4096) CASE(8192)
        default: return -4;
    }
#undef CASE
}

// plan introspection for the tests
extern "C" int spx_emul_plan(int nfft, int* radices /*[5]*/, int* tw_size) {
    const int p = plan_passes(nfft);
    for (int s = 0; s < 5; ++s) radices[s] = plan_radix(nfft, s);
    *tw_size = plan_tw_size(nfft);
    return p;
}
