// spx_tables.h -- host-side construction of the twiddle and window tables (double precision,
// rounded once to float).  Shared by libspx (spx_api.cu) and the CPU re-execution in tests/emul.
#pragma once
#include <math.h>

#include <vector>

#include "spx_fft_core.cuh"

namespace spx {

// layout: for every pass s >= 1, rows t = 1..R-1, columns jm = 0..Ns-1:  W_{Ns*R}^{jm*t}
inline std::vector<float2> build_twiddles(int n) {
    std::vector<float2> tw((size_t)(plan_tw_size(n) > 0 ? plan_tw_size(n) : 1));
    const int passes = plan_passes(n);
    const double two_pi = 6.283185307179586476925286766559;
    for (int s = 1; s < passes; ++s) {
        const int r = plan_radix(n, s), ns = plan_ns(n, s), off = plan_tw_offset(n, s);
        const long long m = (long long)ns * r;
        for (int t = 1; t < r; ++t)
            for (int jm = 0; jm < ns; ++jm) {
                const long long e = ((long long)jm * t) % m;
                const double a = -two_pi * (double)e / (double)m;
                tw[(size_t)off + (size_t)(t - 1) * ns + jm] = make_float2((float)cos(a), (float)sin(a));
            }
    }
    return tw;
}

// symmetric windows exactly as numpy defines them (np.hanning / np.blackman), float64
inline std::vector<double> build_window_f64(int kind, int n) {
    std::vector<double> w((size_t)n, 1.0);
    const double pi = 3.14159265358979323846264338327950288;
    if (n == 1) return w;
    for (int i = 0; i < n; ++i) {
        // numpy: n_ = arange(1 - M, M, 2); hanning = 0.5 + 0.5 cos(pi n_/(M-1));
        //        blackman = 0.42 + 0.5 cos(pi n_/(M-1)) + 0.08 cos(2 pi n_/(M-1))
        const double x = (double)(1 - n + 2 * i);
        if (kind == 1) w[i] = 0.5 + 0.5 * cos(pi * x / (double)(n - 1));
        else if (kind == 2) w[i] = 0.42 + 0.5 * cos(pi * x / (double)(n - 1)) + 0.08 * cos(2.0 * pi * x / (double)(n - 1));
    }
    return w;
}

}  // namespace spx
