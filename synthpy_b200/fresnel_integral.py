"""Wave-optics step after the ray path, with the call shape of the reference's ``src/simulator/fresnel_integral.py``
(SURVEY.md 8f-2): scattered rays -> amplitude / phase grids -> reflect pad + Tukey window -> Fresnel transfer
function -> field at distance z.

What runs where: the interpolation on the triangulation, the padding/window pass, the transfer-function pass and
the crop/scale pass are kernels of csrc/synthpy_b200.cu (per-sample math in csrc/fresnel_core.h); the 2-D FFT is
the library FFT (cuFFT through ``torch.fft``) exactly where the reference calls ``np.fft.fft2``; the Delaunay
triangulation is built on the host with ``scipy.spatial.Delaunay`` -- the Qhull call that the reference's
``LinearNDInterpolator`` makes internally (fresnel_integral.py:71-72) -- once for both interpolated quantities
(the reference triangulates twice).  numpy in -> numpy out, CUDA tensors in -> CUDA tensors out.
"""
import cmath
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import engine
from .engine import _ptr, _stream


def _c128_dev(a):
    t = engine.to_device(a, torch.complex128)
    return t, torch.view_as_real(t)


def scatter_to_grid(px, py, values, x, y, fill_value=0.0, simplices=None):
    """``LinearNDInterpolator((px, py), v, fill_value)(np.meshgrid(x, y))`` for each ``v`` in ``values``
    (fresnel_integral.py:71-77) -> CUDA tensor (len(values), len(y), len(x)).  ``simplices`` (n_tri, 3) may be
    passed to reuse a triangulation; by default it is built here with Qhull on the host."""
    engine.require_cuda()
    px_d, py_d = engine.to_device(px), engine.to_device(py)
    vals = torch.stack([engine.to_device(v) for v in values]).contiguous()
    n = int(px_d.numel())
    if vals.shape[1] != n or py_d.numel() != n:
        raise ValueError("px, py and every value array must have the same length")
    ok = torch.isfinite(px_d) & torch.isfinite(py_d)
    if not bool(ok.all()):              # rays rejected by an aperture are NaN columns: Qhull (and upstream) cannot take them
        if simplices is not None:
            raise ValueError("NaN sample positions with a caller-supplied triangulation")
        px_d, py_d, vals = px_d[ok].contiguous(), py_d[ok].contiguous(), vals[:, ok].contiguous()
        n = int(px_d.numel())
    if n < 3:
        raise ValueError("need at least three finite sample positions to triangulate")
    if simplices is None:
        from scipy.spatial import Delaunay
        pts = np.stack([px_d.cpu().numpy(), py_d.cpu().numpy()], axis=1)
        simplices = Delaunay(pts).simplices
    tri = engine.to_device(np.ascontiguousarray(simplices, dtype=np.int32), torch.int32)
    gx, gy = engine.to_device(x), engine.to_device(y)
    nx, ny = int(gx.numel()), int(gy.numel())
    owner = torch.empty((ny, nx), dtype=torch.int32, device="cuda")
    out = torch.empty((vals.shape[0], ny, nx), dtype=torch.float64, device="cuda")
    L.check(L.lib.sp_scatter_to_grid(_ptr(px_d), _ptr(py_d), _ptr(vals), int(vals.shape[0]), n, _ptr(tri), int(tri.shape[0]),
                                     _ptr(gx), _ptr(gy), nx, ny, float(fill_value), _ptr(owner), _ptr(out), _stream()))
    return out


def bin_to_grid(px, py, amplitudes, phases, x, y):
    """Triangulation-free alternative to ``scatter_to_grid`` for dense ray bundles: U0 on the grid ``x`` x ``y`` as the
    COHERENT MEAN of amp exp(-i phase) over the rays whose nearest node it is (empty nodes: 0).  The sums are taken by
    the library's own detector-binning kernel (``sp_optics_image`` with interferogram accumulation: exact int64 fixed-
    point sums, order-independent), so nothing leaves the GPU and no Qhull call is made -- the host triangulation is
    1.3 s per 3e5 rays against 11 ms for the whole device step.  This is a different estimator
    from upstream's piecewise-linear interpolation (it converges to the same field as the ray density grows; see
    tests/test_gpu_parity.py::test_fresnel_binned_gridding); it is offered as ``propagate(..., gridding='binned')`` and
    is not what parity with the reference is claimed for.  ``x`` and ``y`` must be uniformly spaced.
    Returns (U0 complex128 (len(y), len(x)), rays per node int64)."""
    engine.require_cuda()
    gx, gy = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    dx, dy = (gx[-1] - gx[0]) / (len(gx) - 1), (gy[-1] - gy[0]) / (len(gy) - 1)
    if not (np.allclose(np.diff(gx), dx, rtol=1e-6) and np.allclose(np.diff(gy), dy, rtol=1e-6)):
        raise ValueError("gridding='binned' needs uniformly spaced x and y")
    px_d, py_d = engine.to_device(px), engine.to_device(py)
    amp, ph = engine.to_device(amplitudes), engine.to_device(phases)
    n = int(px_d.numel())
    rf = torch.zeros((4, n), dtype=torch.float64, device="cuda")
    rf[0], rf[2] = px_d, py_d
    jf = torch.empty((2, n), dtype=torch.complex128, device="cuda")
    jf[0] = torch.polar(amp, -ph)                       # amp exp(-i phase): the integrand of U0 (fresnel_integral.py:79)
    jf[1] = 1.0                                         # its sum is the number of rays in the bin
    img = engine.ImageBuffer("interferogram", len(gx), len(gy), (gx[0] - dx / 2, gx[-1] + dx / 2), (gy[0] - dy / 2, gy[-1] + dy / 2))
    engine.optics_image(rf, [], jf=jf, image=img, input_mm=True, want_rays=False)
    planes = img.planes.to(torch.float64) * (2.0 ** -L.PLANE_FRAC_BITS)
    count = planes[2].round()
    u0 = torch.complex(planes[0], planes[1]) / count.clamp(min=1.0)
    return u0, count.to(torch.int64)


def _prepare(a, b, mode, n0, n1, pad_factor, alpha):
    out = torch.empty(((2 * pad_factor + 1) * n0, (2 * pad_factor + 1) * n1), dtype=torch.complex128, device="cuda")
    L.check(L.lib.sp_fresnel_prepare(_ptr(a), _ptr(b), mode, n0, n1, int(pad_factor), float(alpha),
                                     _ptr(torch.view_as_real(out)), _stream()))
    return out


def prepare_field_for_propagation(U0, pad_factor=2, alpha=0.4):
    """fresnel_integral.py:7-24: reflection padding by ``pad_factor`` x size on every side and a Tukey window."""
    as_numpy = not isinstance(U0, torch.Tensor)
    u, ur = _c128_dev(U0)
    out = _prepare(ur, None, 0, int(u.shape[0]), int(u.shape[1]), pad_factor, alpha)
    return out.cpu().numpy() if as_numpy else out


def fresnel_propagate(U0_prepared, L_, wavelength, z, original_shape, pad_factor=2, lanex_fwhm_m=None):
    """fresnel_integral.py:27-59: Fresnel transfer function on the padded grid (sample spacing ``L / original
    size``), optional Gaussian (LANEX) point-spread function, crop back to the original window."""
    as_numpy = not isinstance(U0_prepared, torch.Tensor)
    u, _ = _c128_dev(U0_prepared)
    n0, n1 = (int(v) for v in original_shape)
    m0, m1 = (int(v) for v in u.shape)
    if (m0, m1) != ((2 * pad_factor + 1) * n0, (2 * pad_factor + 1) * n1):
        raise ValueError(f"prepared field {m0}x{m1} is not (2 pad_factor + 1) x the original {n0}x{n1}")
    sigma = 0.0
    if lanex_fwhm_m is not None and lanex_fwhm_m > 0:
        sigma = lanex_fwhm_m / (2 * np.sqrt(2 * np.log(2)))
    spec = torch.fft.fft2(u)                                                    # library FFT (np.fft.fft2 upstream)
    L.check(L.lib.sp_fresnel_transfer(_ptr(torch.view_as_real(spec)), m0, m1, float(L_[0]) / n0, float(L_[1]) / n1,
                                      float(wavelength), float(z), float(sigma), _stream()))
    back = torch.fft.ifft2(spec)
    scale = cmath.exp(1j * (2 * np.pi / wavelength) * z) / (1j * wavelength * z)
    out = torch.empty((n0, n1), dtype=torch.complex128, device="cuda")
    L.check(L.lib.sp_fresnel_finish(_ptr(torch.view_as_real(back)), n0, n1, int(pad_factor), scale.real, scale.imag,
                                    _ptr(torch.view_as_real(out)), _stream()))
    return out.cpu().numpy() if as_numpy else out


def propagate(lwl, x, y, x_length, y_length, jones_vector, amplitudes, phases, z, pad_factor=2, *, return_grids=False,
              gridding="triangulation"):
    """fresnel_integral.py:61-93.  ``jones_vector`` is what upstream passes under that name: the (4, N) ray array
    whose rows 0 and 2 are the sample positions.  Returns the complex field (len(y), len(x)) at distance ``z``.
    ``gridding='triangulation'`` is upstream's LinearNDInterpolator (Qhull on the host, interpolation on the device);
    ``'binned'`` the device-only coherent mean per node (``bin_to_grid``)."""
    as_numpy = not isinstance(jones_vector, torch.Tensor)
    r = engine.to_device(jones_vector)
    if gridding == "binned":
        ok = torch.isfinite(r[0]) & torch.isfinite(r[2])
        u0, _ = bin_to_grid(r[0][ok], r[2][ok], engine.to_device(amplitudes)[ok], engine.to_device(phases)[ok], x, y)
        ny, nx = int(u0.shape[0]), int(u0.shape[1])
        prepared = _prepare(torch.view_as_real(u0.contiguous()), None, 0, ny, nx, pad_factor, 0.4)
        out = fresnel_propagate(prepared, (x_length, y_length), lwl, z, (ny, nx), pad_factor=pad_factor)
        if as_numpy:
            out = out.cpu().numpy()
        return (out, torch.angle(u0).neg(), u0.abs()) if return_grids else out
    if gridding != "triangulation":
        raise ValueError("gridding must be 'triangulation' or 'binned'")
    grids = scatter_to_grid(r[0], r[2], [phases, amplitudes], x, y, fill_value=0.0)
    ny, nx = int(grids.shape[1]), int(grids.shape[2])
    prepared = _prepare(grids[1], grids[0], 1, ny, nx, pad_factor, 0.4)        # U0 = amp exp(-i phase), padded, windowed
    out = fresnel_propagate(prepared, (x_length, y_length), lwl, z, (ny, nx), pad_factor=pad_factor)
    if as_numpy:
        out, grids = out.cpu().numpy(), grids.cpu().numpy()
    return (out, grids[0], grids[1]) if return_grids else out
