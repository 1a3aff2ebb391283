// Microbenchmark: why does the ingest ring lose ~40 us per 16 MB slot against the raw copy ceiling?
// Pipeline per slot k (as spx_ring_commit): H2D(k) on stream A -> event -> kernel(k) (~45 us) on stream B -> event -> D2H(k) on
// stream C.  Modes:
//   0  copies only, both directions at once, no events, no kernel            (the ceiling)
//   1  the ring's dependency structure (events between the three streams)
//   2  as 1, H2D alternating between two streams (the next H2D does not queue behind the previous one's event record)
//   3  as 1, D2H alternating between two streams
//   4  as 1, both alternating
//   5  as 1, but the kernel runs on the H2D stream itself (no A -> B event)
//   6  as 1, enqueued progressively: the host waits for D2H(k - 2) before it submits slot k + 1 (what the ring's producer does)
//   7  as 2, enqueued progressively
//   9  host-driven D2H: the host waits for kernel(k) and only then submits D2H(k) -- no semaphore wait inside the D2H stream;
//      H2D(k + 1) and kernel(k + 1) are submitted before that wait
//   8  as 6 with a real memory-bound kernel (reads + writes 2 x 16 MB for ~45 us) instead of a spinning one
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/copy_gap.cu -o tools/micro/copy_gap
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void spin(long long cycles, unsigned* sink) {
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (sink && threadIdx.x == 12345) *sink = 1;
}

__global__ void stream_kernel(const uint4* in, uint4* out, size_t n, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            uint4 v = in[i];
            v.x += r;
            out[i] = v;
        }
}

int main(int argc, char** argv) {
    const size_t BYTES = (size_t)4 << 22;   // 2^22 ci16 samples
    const int SLOTS = 4, N = 60;
    void *h_in[SLOTS], *h_out[SLOTS], *d_in[SLOTS], *d_out[SLOTS];
    for (int i = 0; i < SLOTS; ++i) {
        cudaHostAlloc(&h_in[i], BYTES, cudaHostAllocPortable);
        cudaHostAlloc(&h_out[i], BYTES, cudaHostAllocPortable);
        cudaMalloc(&d_in[i], BYTES);
        cudaMalloc(&d_out[i], BYTES);
    }
    cudaStream_t sa[2], sb, sc[2];
    for (int i = 0; i < 2; ++i) { cudaStreamCreateWithFlags(&sa[i], cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sc[i], cudaStreamNonBlocking); }
    cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
    cudaEvent_t eh[N], ek[N], ed[N], t0, t1;
    for (int i = 0; i < N; ++i) {
        cudaEventCreateWithFlags(&eh[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ek[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ed[i], cudaEventDisableTiming);
    }
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const long long cyc = (long long)(45e-6 * khz * 1e3);
    for (int mode = 0; mode <= 9; ++mode) {
        const bool progressive = mode >= 6;
        const int base = mode == 6 || mode == 8 ? 1 : (mode == 7 ? 2 : mode);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaDeviceSynchronize();
            cudaEventRecord(t0, sa[0]);
            cudaStreamWaitEvent(sc[0], t0, 0);
            cudaStreamWaitEvent(sc[1], t0, 0);
            cudaStreamWaitEvent(sa[1], t0, 0);
            if (mode == 9) {
                auto submit_up = [&](int k) {
                    const int s = k % SLOTS;
                    cudaMemcpyAsync(d_in[s], h_in[s], BYTES, cudaMemcpyHostToDevice, sa[0]);
                    cudaEventRecord(eh[k], sa[0]);
                    cudaStreamWaitEvent(sb, eh[k], 0);
                    spin<<<1, 32, 0, sb>>>(cyc, nullptr);
                    cudaEventRecord(ek[k], sb);
                };
                submit_up(0);
                for (int k = 0; k < N; ++k) {
                    if (k + 1 < N) {
                        if (k + 1 >= SLOTS) cudaEventSynchronize(ed[k + 1 - SLOTS]);   // slot reuse
                        submit_up(k + 1);
                    }
                    cudaEventSynchronize(ek[k]);
                    cudaMemcpyAsync(h_out[k % SLOTS], d_out[k % SLOTS], BYTES, cudaMemcpyDeviceToHost, sc[0]);
                    cudaEventRecord(ed[k], sc[0]);
                }
            } else
            for (int k = 0; k < N; ++k) {
                const int s = k % SLOTS;
                cudaStream_t A = sa[(base == 2 || base == 4) ? (k & 1) : 0], Cs = sc[(base == 3 || base == 4) ? (k & 1) : 0];
                if (progressive && k >= 3) cudaEventSynchronize(ed[k - 3]);
                if (k >= SLOTS && base != 0 && !progressive) cudaStreamWaitEvent(A, ed[k - SLOTS], 0);   // slot reuse (the host's collect / release)
                cudaMemcpyAsync(d_in[s], h_in[s], BYTES, cudaMemcpyHostToDevice, A);
                if (base == 0) {
                    cudaMemcpyAsync(h_out[s], d_out[s], BYTES, cudaMemcpyDeviceToHost, Cs);
                    continue;
                }
                if (base == 5) {
                    spin<<<1, 32, 0, A>>>(cyc, nullptr);
                    cudaEventRecord(ek[k], A);
                } else {
                    cudaEventRecord(eh[k], A);
                    cudaStreamWaitEvent(sb, eh[k], 0);
                    if (mode == 8) stream_kernel<<<148 * 8, 256, 0, sb>>>((const uint4*)d_in[s], (uint4*)d_out[s], BYTES / 16, 10);
                    else spin<<<1, 32, 0, sb>>>(cyc, nullptr);
                    cudaEventRecord(ek[k], sb);
                }
                cudaStreamWaitEvent(Cs, ek[k], 0);
                cudaMemcpyAsync(h_out[s], d_out[s], BYTES, cudaMemcpyDeviceToHost, Cs);
                cudaEventRecord(ed[k], Cs);
            }
            cudaStreamWaitEvent(sa[0], ed[N - 1], 0);
            if (mode == 0) { cudaEventRecord(ed[0], sc[0]); cudaStreamWaitEvent(sa[0], ed[0], 0); cudaEventRecord(ed[1], sc[1]); cudaStreamWaitEvent(sa[0], ed[1], 0); cudaEventRecord(ed[2], sa[1]); cudaStreamWaitEvent(sa[0], ed[2], 0); }
            if (N > 1) cudaStreamWaitEvent(sa[0], ed[N - 2], 0);
            cudaEventRecord(t1, sa[0]);
            cudaEventSynchronize(t1);
            float ms;
            cudaEventElapsedTime(&ms, t0, t1);
            if (ms < best) best = ms;
        }
        printf("{\"mode\": %d, \"us_per_slot\": %.1f, \"gbs_each_way\": %.2f, \"err\": \"%s\"}\n", mode, best * 1e3 / N, BYTES * (double)N / (best * 1e-3) / 1e9,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
