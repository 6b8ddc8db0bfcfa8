// Per-sample math of the wave-optics step that follows the ray path: scattered rays ->
// detector grid on a triangulation, reflect padding + Tukey window, Fresnel transfer function.
// Host + device (SP_HD) like ray_core.h: inlined into the kernels of synthpy_b200.cu and compiled by g++ into
// tests/host_harness.cpp for the CPU-side checks.  The 2-D FFT between `prepare` and `finish` is the library FFT
// (cuFFT through torch.fft) -- the reference calls np.fft.fft2 at the same place.
//
// Reference semantics restated here (src/simulator/fresnel_integral.py):
//   bary2 / tri_value : scipy.interpolate.LinearNDInterpolator on its Delaunay triangulation   (:71-77)
//   reflect_idx       : np.pad(mode='reflect')                                                 (:15)
//   tukey_w           : scipy.signal.windows.tukey(M, alpha), symmetric                        (:17-20)
//   fft_freq          : np.fft.fftfreq                                                         (:36-37)
//   prepare_sample    : U0 = amp exp(-i phase); pad; window                                    (:79-84, :7-24)
//   transfer_sample   : exp(-i pi lambda z (f0^2 + f1^2)) [x Gaussian PSF]                     (:41-49)
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef SP_HD
#if defined(__CUDACC__)
#define SP_HD __host__ __device__ __forceinline__
#else
#define SP_HD inline
#endif
#endif

namespace sp {

#define SP_PI 3.141592653589793238462643383279502884

// np.pad 'reflect': period 2(n-1), the edge sample is not repeated; valid for any distance from the array.
SP_HD long long reflect_idx(long long i, long long n) {
    if (n == 1) return 0;
    const long long p = 2 * (n - 1);
    long long m = i % p;
    if (m < 0) m += p;
    return m < n ? m : p - m;
}

// scipy.signal.windows.tukey(M, alpha) sample i (sym=True).
SP_HD double tukey_w(long long i, long long M, double alpha) {
    if (M == 1 || alpha <= 0.0) return 1.0;
    if (alpha >= 1.0) return 0.5 - 0.5 * cos(2.0 * SP_PI * (double)i / (double)(M - 1));
    const long long width = (long long)floor(alpha * (double)(M - 1) / 2.0);
    if (i <= width) return 0.5 * (1.0 + cos(SP_PI * (-1.0 + 2.0 * (double)i / alpha / (double)(M - 1))));
    if (i >= M - width - 1) return 0.5 * (1.0 + cos(SP_PI * (-2.0 / alpha + 1.0 + 2.0 * (double)i / alpha / (double)(M - 1))));
    return 1.0;
}

// np.fft.fftfreq(m, d)[k] = k_signed * (1 / (m d))
SP_HD double fft_freq(long long k, long long m, double d) {
    const long long ks = (k < (m + 1) / 2) ? k : k - m;
    return (double)ks * (1.0 / ((double)m * d));
}

// U0 = amp exp(-i phase) (fresnel_integral.py:79)
SP_HD void u0_from_amp_phase(double amp, double phase, double& re, double& im) {
    double sn, cs;
    sincos(phase, &sn, &cs);
    re = amp * cs;
    im = -(amp * sn);
}

// One sample (i0, i1) of the padded, windowed field.  mode 0: a = interleaved complex U0; mode 1: a = amplitude,
// b = phase.  Source arrays are [n0][n1]; the padded grid is [(2 pad + 1) n0][(2 pad + 1) n1].  (The kernels compute
// the same factors once per row / column / source sample instead of once per padded sample.)
SP_HD void prepare_sample(const double* a, const double* b, int mode, long long n0, long long n1, long long pad,
                          double alpha, long long i0, long long i1, double& re, double& im) {
    const long long s0 = reflect_idx(i0 - pad * n0, n0), s1 = reflect_idx(i1 - pad * n1, n1);
    const long long src = s0 * n1 + s1;
    double ur, ui;
    if (mode == 0) {
        ur = a[2 * src];
        ui = a[2 * src + 1];
    } else {
        u0_from_amp_phase(a[src], b[src], ur, ui);
    }
    const double w = tukey_w(i0, (2 * pad + 1) * n0, alpha) * tukey_w(i1, (2 * pad + 1) * n1, alpha);
    re = ur * w;
    im = ui * w;
}

// Multiply one spectrum sample by the Fresnel transfer function (and the optional Gaussian PSF, sigma > 0);
// f0, f1 are the sample's frequencies fft_freq(k, m, d).
SP_HD void transfer_apply(double& re, double& im, double f0, double f1, double wavelength, double z, double sigma) {
    const double F2 = f0 * f0 + f1 * f1;
    double sn, cs;
    sincos(-(SP_PI * wavelength * z * F2), &sn, &cs);
    double hr = cs, hi = sn;
    if (sigma > 0.0) {
        const double g = exp(-2.0 * (SP_PI * sigma) * (SP_PI * sigma) * F2);
        hr *= g;
        hi *= g;
    }
    const double r = re * hr - im * hi, i = re * hi + im * hr;
    re = r;
    im = i;
}

SP_HD void transfer_sample(double& re, double& im, long long k0, long long k1, long long m0, long long m1, double d0,
                           double d1, double wavelength, double z, double sigma) {
    transfer_apply(re, im, fft_freq(k0, m0, d0), fft_freq(k1, m1, d1), wavelength, z, sigma);
}

// Barycentric coordinates of p in triangle (a, b, c); false for a degenerate triangle.
SP_HD bool bary2(double ax, double ay, double bx, double by, double cx, double cy, double px, double py, double& l0,
                 double& l1, double& l2) {
    const double d = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay);
    if (d == 0.0 || !(d == d)) return false;
    // same form as scipy's transform: [l0, l1] = T (p - c), l2 = 1 - l0 - l1 with c the LAST vertex
    const double qx = px - cx, qy = py - cy;
    const double dd = (ax - cx) * (by - cy) - (bx - cx) * (ay - cy);
    l0 = ((by - cy) * qx - (bx - cx) * qy) / dd;
    l1 = (-(ay - cy) * qx + (ax - cx) * qy) / dd;
    l2 = 1.0 - l0 - l1;
    return true;
}

#define SP_TRI_EPS 2.220446049250313e-14 /* 100 DBL_EPSILON: the inside tolerance of scipy's find_simplex */

SP_HD bool tri_inside(double l0, double l1, double l2) {
    return l0 >= -SP_TRI_EPS && l1 >= -SP_TRI_EPS && l2 >= -SP_TRI_EPS;
}

// first index k in [0, n) with g[k] >= v (n if none); g ascending
SP_HD int lower_bound_d(const double* g, int n, double v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (g[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace sp
