#!/usr/bin/env python
"""Device timing of the wave-optics step at detector scale: N scattered rays -> (ny, nx) grids ->
pad x5 + Tukey window -> FFT -> Fresnel transfer -> inverse FFT -> crop.  CUDA events per stage, algorithmic bytes
per stage against the measured HBM peak.  One JSON line on stdout.

    python examples/fresnel_bench.py [--rays 300000] [--nx 1724] [--ny 1287] [--reps 5]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

from synthpy_b200 import _lib as L, engine, fresnel_integral as FI  # noqa: E402
from synthpy_b200.engine import _ptr, _stream  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=float, default=3e5)
    ap.add_argument("--nx", type=int, default=3448 // 2)
    ap.add_argument("--ny", type=int, default=2574 // 2)
    ap.add_argument("--pad", type=int, default=2)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    engine.require_cuda()
    n, nx, ny, pad = int(a.rays), a.nx, a.ny, a.pad
    rng = np.random.default_rng(0)
    Lx, Ly, lwl, z = 18e-3, 13.5e-3, 1064e-9, 0.3
    px, py = rng.uniform(-0.49 * Lx, 0.49 * Lx, n), rng.uniform(-0.49 * Ly, 0.49 * Ly, n)
    amp = 1.0 + 0.3 * np.cos(2 * np.pi * px / 4e-3)
    phase = 40.0 * np.exp(-(px ** 2 + py ** 2) / (3e-3) ** 2)
    x, y = np.linspace(-Lx / 2, Lx / 2, nx), np.linspace(-Ly / 2, Ly / 2, ny)
    from scipy.spatial import Delaunay
    t0 = time.perf_counter()
    tri = Delaunay(np.stack([px, py], 1)).simplices
    t_qhull = time.perf_counter() - t0

    pxd, pyd = engine.to_device(px), engine.to_device(py)
    vals = torch.stack([engine.to_device(phase), engine.to_device(amp)]).contiguous()
    trid = engine.to_device(np.ascontiguousarray(tri, dtype=np.int32), torch.int32)
    gx, gy = engine.to_device(x), engine.to_device(y)
    owner = torch.empty((ny, nx), dtype=torch.int32, device="cuda")
    grids = torch.empty((2, ny, nx), dtype=torch.float64, device="cuda")
    m0, m1 = (2 * pad + 1) * ny, (2 * pad + 1) * nx
    padded = torch.empty((m0, m1), dtype=torch.complex128, device="cuda")
    out = torch.empty((ny, nx), dtype=torch.complex128, device="cuda")

    def k_scatter():
        L.check(L.lib.sp_scatter_to_grid(_ptr(pxd), _ptr(pyd), _ptr(vals), 2, n, _ptr(trid), int(trid.shape[0]), _ptr(gx), _ptr(gy),
                                         nx, ny, 0.0, _ptr(owner), _ptr(grids), _stream()))

    def k_prepare():
        L.check(L.lib.sp_fresnel_prepare(_ptr(grids[1]), _ptr(grids[0]), 1, ny, nx, pad, 0.4, _ptr(torch.view_as_real(padded)), _stream()))

    def k_transfer():
        L.check(L.lib.sp_fresnel_transfer(_ptr(torch.view_as_real(padded)), m0, m1, Lx / ny, Ly / nx, lwl, z, 0.0, _stream()))

    def k_finish():
        L.check(L.lib.sp_fresnel_finish(_ptr(torch.view_as_real(padded)), ny, nx, pad, 1.0, 0.0, _ptr(torch.view_as_real(out)), _stream()))

    ms = {"scatter_to_grid": timed(k_scatter, a.reps), "prepare": timed(k_prepare, a.reps), "transfer": timed(k_transfer, a.reps),
          "finish": timed(k_finish, a.reps), "cufft_fft2+ifft2": timed(lambda: torch.fft.ifft2(torch.fft.fft2(padded)), max(1, a.reps // 2))}
    ms["propagate_end_to_end_excl_qhull"] = timed(
        lambda: FI.fresnel_propagate(FI._prepare(grids[1], grids[0], 1, ny, nx, pad, 0.4), (Lx, Ly), lwl, z, (ny, nx), pad_factor=pad), 2)
    npix, npad = nx * ny, m0 * m1
    alg = {"scatter_to_grid": npix * (4 + 4 + 16) + trid.numel() * 4 + 3 * trid.shape[0] * 16,   # owner fill+read, 2 grids out, triangles + vertices
           "prepare": npad * 16 + npix * (16 + 32),      # u0 pass (16 in, 16 out) + gather + padded store
           "transfer": npad * 32, "finish": npix * 32}
    peak = 6556.2
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    roof = {k: {"ms": ms[k], "algorithmic_bytes": int(alg[k]), "GBps": alg[k] / ms[k] / 1e6, "frac_of_hbm_peak": alg[k] / ms[k] / 1e6 / peak}
            for k in alg}
    print(json.dumps({"bench": "fresnel_step", "rays": n, "grid": [ny, nx], "padded": [m0, m1], "triangles": int(trid.shape[0]),
                      "host_qhull_s": t_qhull, "ms": ms, "roofline": roof, "hbm_peak_GBps": peak}))


if __name__ == "__main__":
    main()
