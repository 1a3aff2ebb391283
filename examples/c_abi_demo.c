/* c_abi_demo.c -- libspx from plain C: the three hot lines of the reference stream loop
 * (/root/reference/app/sdr/streamer.py:119-121) for one rx buffer, then a Welch block of an int16 stream.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lsdr_iq_visualizer_b200/csrc -lspx -lm -o c_abi_demo
 *   LD_LIBRARY_PATH=sdr_iq_visualizer_b200/csrc ./c_abi_demo
 *
 * Exits 0 and prints "c_abi_demo ok" when the results match a direct DFT of the same samples; exits 3 when no
 * CUDA device is present (the library has no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "spx.h"

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc_ = (call);                                                    \
        if (rc_ != SPX_OK) {                                                 \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, spx_last_error()); \
            return rc_ == SPX_E_NODEVICE ? 3 : 1;                            \
        }                                                                    \
    } while (0)

int main(void) {
    enum { N = 64 };
    const double two_pi = 6.283185307179586;
    int ndev = 0;
    int rc = spx_device_count(&ndev);
    if (rc != SPX_OK || ndev == 0) {
        fprintf(stderr, "no CUDA device: %s\n", spx_last_error());
        return 3;
    }
    if (spx_abi_version() != SPX_ABI_VERSION) return 1;

    /* one rx buffer: tone on bin 5 plus a constant */
    float x[2 * N];
    for (int n = 0; n < N; ++n) {
        x[2 * n] = (float)(100.0 * cos(two_pi * 5.0 * n / N) + 3.0);
        x[2 * n + 1] = (float)(100.0 * sin(two_pi * 5.0 * n / N));
    }
    spx_plan_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.nfft = N; cfg.hop = N; cfg.window = SPX_WINDOW_RECT; cfg.in_fmt = SPX_FMT_CF32;
    cfg.in_scale = 1.0f; cfg.db_eps = 1e-12f;
    spx_plan* plan = NULL;
    CHECK(spx_plan_create(&plan, &cfg));

    float db[N];
    spx_stft_args a;
    memset(&a, 0, sizeof(a));
    a.struct_size = sizeof(a);
    a.mem = SPX_MEM_HOST;
    a.in = x; a.n_samples = N; a.n_streams = 1;
    a.db_rows = db;
    CHECK(spx_stft_exec(plan, &a));
    if (a.n_frames_out != 1) return 1;

    /* reference: 20*log10(|fftshift(fft(x))| + 1e-12) by direct summation in double */
    double worst = 0.0;
    for (int j = 0; j < N; ++j) {
        const int k = (j + N / 2) % N;
        double re = 0.0, im = 0.0;
        for (int n = 0; n < N; ++n) {
            const double c = cos(two_pi * k * n / N), s = -sin(two_pi * k * n / N);
            re += x[2 * n] * c - x[2 * n + 1] * s;
            im += x[2 * n] * s + x[2 * n + 1] * c;
        }
        const double ref = 20.0 * log10(sqrt(re * re + im * im) + 1e-12);
        if (ref > 0.0 && fabs(ref - db[j]) > worst) worst = fabs(ref - db[j]);   /* the two occupied bins */
    }
    if (worst > 1e-3) { fprintf(stderr, "dB mismatch %g\n", worst); return 1; }
    CHECK(spx_plan_destroy(plan));

    /* an int16 stream: 4096-point Hann, 75 %% overlap, Welch block + classifier measurements */
    const int NF = 4096, HOP = 1024;
    const long L = 1 << 18;
    short* iq = (short*)malloc(sizeof(short) * 2 * (size_t)L);
    for (long n = 0; n < L; ++n) {
        iq[2 * n] = (short)lrint(1000.0 * cos(two_pi * 0.125 * (double)n));
        iq[2 * n + 1] = (short)lrint(1000.0 * sin(two_pi * 0.125 * (double)n));
    }
    cfg.nfft = NF; cfg.hop = HOP; cfg.window = SPX_WINDOW_HANN; cfg.in_fmt = SPX_FMT_CI16;
    CHECK(spx_plan_create(&plan, &cfg));
    double* welch = (double*)calloc(NF, sizeof(double));
    double* pxx_db = (double*)calloc(NF, sizeof(double));
    memset(&a, 0, sizeof(a));
    a.struct_size = sizeof(a);
    a.mem = SPX_MEM_HOST;
    a.in = iq; a.n_samples = L; a.n_streams = 1;
    a.welch_acc = welch;
    CHECK(spx_stft_exec(plan, &a));
    if (a.n_frames_out != spx_frame_count(L, NF, HOP)) return 1;
    CHECK(spx_welch_finalize(plan, SPX_MEM_HOST, welch, a.n_frames_out, 61.44e6, NULL, pxx_db, NULL));
    spx_features f;
    CHECK(spx_classify_features(0, SPX_MEM_HOST, pxx_db, 1, NF, 1, NF, &f, NULL, 0, NULL, NULL));
    /* +0.125 cycles/sample -> bin N/8 above DC -> position N/2 + N/8 in fftshift order */
    if (f.argmax != NF / 2 + NF / 8 || !(f.snr_db > 40.0)) { fprintf(stderr, "argmax %d snr %g\n", f.argmax, f.snr_db); return 1; }
    CHECK(spx_plan_destroy(plan));
    free(iq); free(welch); free(pxx_db);
    printf("c_abi_demo ok (max dB error %.2e, tone at bin %d, SNR %.1f dB)\n", worst, f.argmax, f.snr_db);
    return 0;
}
