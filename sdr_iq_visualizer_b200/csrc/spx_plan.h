// spx_plan.h -- the opaque plan object behind include/spx.h (host-side only).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <vector>

#include "spx_internal.h"

namespace spx {

struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);  // grow-only
    void release();
};

}  // namespace spx

struct spx_plan {
    spx_plan_config cfg{};
    int sm_count = 0;
    float* d_win = nullptr;   // window * in_scale (nullptr: rect and scale 1)
    float2* d_tw = nullptr;   // twiddle table of the shared-memory kernel (nfft <= 8192)
    std::vector<double> win64;
    double sum_w2 = 0.0, sum_w = 0.0;
    cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    // the ingest ring uploads consecutive slots on alternating streams (s_h2d, s_h2d_alt): two uploads in flight keep the H2D
    // direction full while the D2H direction is busy -- a single copy at a time leaves ~35 us per 16 MB slot on the table
    // (tools/micro/copy_gap.cu: 386 vs 355 us per slot; needs >= 5 slots in flight in the ring to show, profiles/r02_e2e_ring_depth.jsonl)
    cudaStream_t s_h2d_alt = nullptr;
    std::vector<cudaEvent_t> events;
    size_t piece_bytes = 16u << 20;  // H2D piece size of the host pipeline
    size_t peer_piece_bytes = 48u << 20;  // uint8 rows per piece of the peer-output pipeline (two staging buffers)
    // staging for SPX_MEM_HOST execution (grow-only)
    spx::DevBuf st_in, st_db, st_wf, st_spec, st_welch, st_max, st_misc, st_flush;
    // large-N (four-step) path, nfft >= 16384
    int big_n1 = 0, big_n2 = 0;
    float2* d_big_tw = nullptr;  // one allocation holding the four tables below
    float2 *d_tw1 = nullptr, *d_tw2 = nullptr, *d_wn_fine = nullptr, *d_wn_coarse = nullptr;
    size_t big_scratch_bytes = 384u << 20;  // two halves of 192 MB: frames per batch = half / (nfft * 8); measured best on B200
    spx::DevBuf st_big;
    float2* d_big2 = nullptr;      // K2v2 (single-kernel 65536-point path): N = 4096 twiddle table followed by the base table
    size_t big2_tw_count = 0;
    float* d_big2_win = nullptr;   // K2v2 window table in phase A's thread order
    cudaStream_t s_big_aux = nullptr;             // column kernels of batch k+1 run here, next to the row kernels of batch k
    cudaEvent_t ev_big[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // entry, A done x2, B done x2
    // Bluestein path (nfft not a power of two, or < 16): inner power-of-two plan of length blu_m
    int blu_m = 0;
    spx_plan* blu_inner = nullptr;
    float2* d_blu = nullptr;  // [N] window*scale*chirp, [N] chirp/M, [M] FFT_M(conj chirp) in fftshift order
    // ordering of plan-owned scratch between calls on different streams
    cudaEvent_t ev_scratch = nullptr;
    cudaStream_t scratch_stream = nullptr;
    bool scratch_stream_valid = false;
    std::mutex mu;
};

namespace spx {
int plan_event(spx_plan* pl, size_t i, cudaEvent_t* out);
int stft_launch_device(spx_plan* pl, const void* in, long long n_streams, long long stream_stride, long long frames,
                       float* db_rows, unsigned char* wf_rows, float2* spec_rows, double* welch_acc, float* maxhold,
                       float vmin, float vmax, cudaStream_t st, int sys_atomics = 0);
int bigfft_plan_init(spx_plan* pl);
int bigfft_launch_stream(spx_plan* pl, const void* in, long long frames, long long row0, float* db_rows,
                         unsigned char* wf_rows, float2* spec_rows, double* welch_acc, float* maxhold, float vmin,
                         float vmax, cudaStream_t st, int sys_atomics = 0);
int peer_reduce_launch(const double* w_local, const float* m_local, double* w_peer, float* m_peer, long long n, cudaStream_t st);
int big2_plan_init(spx_plan* pl);
bool big2_eligible(const spx_plan* pl, const void* in, const float* db_rows, const float2* spec_rows);
int big2_launch_stream(spx_plan* pl, const void* in, long long frames, long long row0, unsigned char* wf_rows, double* welch_acc,
                       float* maxhold, float vmin, float vmax, cudaStream_t st, int sys_atomics);
int welch_finalize_launch(const double* acc, int n, double inv_norm, double* pxx, double* pxx_db, cudaStream_t st);
int bluestein_plan_init(spx_plan* pl);
int bluestein_launch_stream(spx_plan* pl, const void* in, long long frames, long long row0, float* db_rows,
                            unsigned char* wf_rows, float2* spec_rows, double* welch_acc, float* maxhold, float vmin,
                            float vmax, cudaStream_t st, int sys_atomics = 0);
}  // namespace spx
