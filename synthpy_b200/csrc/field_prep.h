// Field preparation shared by the CUDA kernels and the CPU self-test build: axis tables, np.gradient
// coefficients in float32 and the per-cell pack function.
//
// Reference: ScalarDomain.calc_dndr, src/solvers-legacy/full_solver.py:211-234
//   ne_nc = float32(ne / nc);  dnd{x,y,z} = -0.5 c^2 * np.gradient(ne_nc, axis_a, axis=a)   (all float32)
// and n_refrac, full_solver.py:236-239,270-274.  np.gradient = numpy/lib/_function_base_impl.py::gradient with
// edge_order=1: second-order interior stencil for NON-uniform spacing (the float32-rounded linspace axes
// are not equispaced), uniform stencil when every float32 spacing is equal, first-order edges.
#pragma once
#include <vector>

#include "ray_core.h"

namespace sp {

// never-contracted float32 arithmetic (NumPy evaluates each ufunc separately: no FMA)
SP_HD float fmul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
SP_HD float fadd_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
SP_HD float fsub_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
SP_HD float fdiv_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}

// np.gradient coefficients of one axis, evaluated in float32 exactly as NumPy does.
struct AxisCoef {
    int n, uniform;
    float two_dx, dx0, dxn;
    std::vector<float> a, b, c;   // interior i = 1..n-2 stored at [i]
};

inline AxisCoef axis_coef(const float* g, int n) {
    AxisCoef A; A.n = n;
    std::vector<float> dx(n - 1);
    for (int i = 0; i < n - 1; ++i) dx[i] = fsub_rn(g[i + 1], g[i]);               // np.diff
    A.uniform = 1;
    for (int i = 1; i < n - 1; ++i) if (dx[i] != dx[0]) A.uniform = 0;             // (diffx == diffx[0]).all()
    A.a.assign(n, 0.f); A.b.assign(n, 0.f); A.c.assign(n, 0.f);
    if (A.uniform) {
        A.two_dx = fmul_rn(2.0f, dx[0]); A.dx0 = dx[0]; A.dxn = dx[0];
    } else {
        A.two_dx = 0.f; A.dx0 = dx[0]; A.dxn = dx[n - 2];
        for (int i = 1; i < n - 1; ++i) {
            const float d1 = dx[i - 1], d2 = dx[i];
            const float s = fadd_rn(d1, d2);
            A.a[i] = fdiv_rn(-d2, fmul_rn(d1, s));               // a = -(dx2)/(dx1 * (dx1 + dx2))
            A.b[i] = fdiv_rn(fsub_rn(d2, d1), fmul_rn(d1, d2));  // b = (dx2 - dx1) / (dx1 * dx2)
            A.c[i] = fdiv_rn(d1, fmul_rn(d2, s));                // c = dx1 / (dx2 * (dx1 + dx2))
        }
    }
    return A;
}

// Interpolation tables of one axis: {g[i], 1/(g[i+1]-g[i])} in float64 and float32 + first-guess constants.
struct AxisTables {
    std::vector<d2> t64; std::vector<f2> t32;
    double g0, inv_d, lo, hi;
};

inline bool build_axis_tables(const float* g, int n, AxisTables& T) {
    for (int i = 0; i + 1 < n; ++i) if (!(g[i + 1] > g[i])) return false;
    T.t64.resize(n); T.t32.resize(n);
    for (int i = 0; i < n; ++i) {
        T.t64[i].x = (double)g[i]; T.t32[i].x = g[i];
        if (i + 1 < n) {
            T.t64[i].y = 1.0 / ((double)g[i + 1] - (double)g[i]);
            T.t32[i].y = fdiv_rn(1.0f, fsub_rn(g[i + 1], g[i]));
        } else { T.t64[i].y = 0.0; T.t32[i].y = 0.f; }
    }
    T.g0 = g[0]; T.lo = g[0]; T.hi = g[n - 1];
    T.inv_d = (double)(n - 1) / ((double)g[n - 1] - (double)g[0]);
    return true;
}

struct StencilDev {          // per caller axis; a/b/c live in device (or host, for the self-test) memory
    const float *a, *b, *c;
    float two_dx, dx0, dxn;
    int uniform, n;
};

struct PackArgs {
    int n[3];      // caller dims nx, ny, nz
    int perm[3];   // kernel axis -> caller axis
    int nk[3];     // kernel-frame dims
    StencilDev st[3];
    float k32;     // float32(-0.5 c^2)
    double omega;
    int flags;     // SP_FIELD_PHASE | SP_FIELD_PHASE_F64
};

SP_HD float normalise_ne(double ne, double nc) { return (float)(ne / nc); }                  // float32(ne / nc)
SP_HD float normalise_ne(float ne, double nc) { return fdiv_rn(ne, (float)nc); }            // f32 array / python float

SP_HD float grad1(const float* f, long long idx, long long stride, int i, const StencilDev& S) {
    if (i == 0) return fdiv_rn(fsub_rn(f[idx + stride], f[idx]), S.dx0);
    if (i == S.n - 1) return fdiv_rn(fsub_rn(f[idx], f[idx - stride]), S.dxn);
    if (S.uniform) return fdiv_rn(fsub_rn(f[idx + stride], f[idx - stride]), S.two_dx);
    return fadd_rn(fadd_rn(fmul_rn(S.a[i], f[idx - stride]), fmul_rn(S.b[i], f[idx])), fmul_rn(S.c[i], f[idx + stride]));
}

// Packed-cell index t (kernel frame, w fastest) -> caller-frame flat index [x][y][z] and per-axis indices.
SP_HD long long unpack_index(long long t, const PackArgs& P, int ic[3]) {
    int ik[3];
    ik[2] = (int)(t % P.nk[2]);
    const long long r = t / P.nk[2];
    ik[1] = (int)(r % P.nk[1]);
    ik[0] = (int)(r / P.nk[1]);
    ic[P.perm[0]] = ik[0]; ic[P.perm[1]] = ik[1]; ic[P.perm[2]] = ik[2];
    return ((long long)ic[0] * P.n[1] + ic[1]) * P.n[2] + ic[2];
}

// Where packed cell t (kernel frame, w fastest) lives inside the float4 array: t itself in the row layout, its slot in
// the 8 x 8 x 8 brick grid in the A/B layout (-DSP_BRICK).
SP_HD long long packed_slot(long long t, const PackArgs& P) {
#ifdef SP_BRICK
    const int iw = (int)(t % P.nk[2]);
    const long long r = t / P.nk[2];
    const int iv = (int)(r % P.nk[1]), iu = (int)(r / P.nk[1]);
    const long long nbv = (P.nk[1] + 7) >> 3, nbw = (P.nk[2] + 7) >> 3;
    return (((long long)(iu >> 3) * nbv + (iv >> 3)) * nbw + (iw >> 3)) * 512 + ((iu & 7) << 6) + ((iv & 7) << 3) + (iw & 7);
#else
    (void)P;
    return t;
#endif
}

// n - 1 with n = sqrt(1 - (5.64e4 sqrt(ne 1e-6) / omega)^2)      (full_solver.py:236-239,270-274)
SP_HD double refr_minus_one(double ne, double omega) {
    const double ope = mul_rn(5.64e4, sqrt(mul_rn(ne, 1e-6)));
    const double q = ope / omega;
    return sqrt(add_rn(1.0, -mul_rn(q, q))) - 1.0;
}

template <typename NE>
SP_HD f4 pack_cell(long long t, const float* ne_nc, const NE* ne, const PackArgs& P, double& nm1) {
    int ic[3];
    const long long idx = unpack_index(t, P, ic);
    const long long sx = (long long)P.n[1] * P.n[2], sy = P.n[2];
    float g[3];
    g[0] = fmul_rn(P.k32, grad1(ne_nc, idx, sx, ic[0], P.st[0]));
    g[1] = fmul_rn(P.k32, grad1(ne_nc, idx, sy, ic[1], P.st[1]));
    g[2] = fmul_rn(P.k32, grad1(ne_nc, idx, 1, ic[2], P.st[2]));
    f4 v; v.x = g[P.perm[0]]; v.y = g[P.perm[1]]; v.z = g[P.perm[2]]; v.w = 0.f;
    nm1 = 0.0;
    if (P.flags & 3) {
        nm1 = refr_minus_one((double)ne[idx], P.omega);
        v.w = (float)nm1;
    }
    return v;
}

}  // namespace sp
