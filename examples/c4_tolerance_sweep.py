"""BASELINE config C4: refractometry + knife-edge schlieren with the adaptive RK45, tolerance sweep.

For each (rtol, atol) of the sweep -- SciPy's defaults (what full_solver.py:391 runs with), two intermediate
settings and the tolerances the reference's diffrax variants intend (1e-7 / 1e-9) plus one tighter -- trace the same
device-generated rays through the 512^3 turbulent field and report attempted steps per ray, rays*steps/s and the L1
distance of each detector image to the image of the tightest setting (normalised by the image's total counts).

    python examples/c4_tolerance_sweep.py [--rays 2000000] [--grid 512] [--bundle] [--out sweep.jsonl]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

import bench
from synthpy_b200 import beam as B, diagnostics as D, domain as Dm, engine, propagator as P

SWEEP = [(1e-3, 1e-6), (1e-5, 1e-8), (1e-7, 1e-9), (1e-9, 1e-12)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=float, default=2e6)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--bundle", action="store_true", help="one step size per 32-ray bundle instead of per ray")
    ap.add_argument("--bin-scale", type=int, default=4)
    ap.add_argument("--max-steps", type=int, default=200000,
                    help="attempt cap per ray: below the float64 noise floor (rtol 1e-9) a few rays collapse to the minimum step, "
                         "exactly as solve_ivp would; the cap bounds the time the rest of their warp waits")
    ap.add_argument("--reps", type=int, default=2, help="passes per setting (the last one is timed); 1 for the 1e8-ray runs, "
                    "whose kernels last seconds to minutes and need no warm-up")
    ap.add_argument("--tolerances", default=None, help="comma-separated subset of sweep indices 0..3")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    n = int(a.rays)
    ne = bench.build_ne(bench.parse(["--workload", "C4", "--grid", str(a.grid)]), "cuda")
    dom = Dm.ScalarDomain(bench.LENGTHS, a.grid)
    dom.external_ne(ne)
    dom.device_field(bench.LWL)
    del ne
    beam = B.Beam(n, bench.BEAM_R, bench.BEAM_DIV, bench.EXTENT, device=True, seed=2, beam_type="circular")
    rows, images = [], []
    sweep = SWEEP if a.tolerances is None else [SWEEP[int(i)] for i in a.tolerances.split(",")]
    for rtol, atol in [(None, None)] + sweep:              # first row: the fixed-step production mode, for comparison
        specs = [D.spec("refracto_incoherent", bin_scale=a.bin_scale),
                 D.spec("schlieren_knife", bin_scale=a.bin_scale, offset=0.1, axis=2, direction=1)]
        if rtol is None:
            kw = dict(lwl=bench.LWL, method="rk4", ds=0.5 * dom.cell_size())
        else:
            kw = dict(lwl=bench.LWL, method="rk45_bundle" if a.bundle else "rk45", rtol=rtol, atol=atol, max_steps=a.max_steps)
        for rep in range(max(1, a.reps)):                 # earlier passes warm caches / clocks, the last is timed
            for s in specs:
                s.image.zero_()
            engine.propagate_kernel_ms()
            st, _ = P.solve_and_image(dom, beam, bench.EXTENT, specs, n_rays=n, sync=True, **kw)
        ms, _ = engine.propagate_kernel_ms()
        sd = st                                             # sync=True returns the counters as a dict
        images.append([s.image.result().double().clone() for s in specs])
        rows.append({"rtol": rtol, "atol": atol, "mode": "rk4, ds = 0.5 cell" if rtol is None else ("bundle" if a.bundle else "per ray"), "rays": n,
                     "steps_per_ray": sd["ray_steps"] / n, "accepted_per_ray": sd["ray_steps_acc"] / n,
                     "rays_capped": sd["rays_capped"], "kernel_ms": ms, "rays_steps_per_s": sd["ray_steps"] / (ms * 1e-3),
                     "rays_per_s": n / (ms * 1e-3), "rays_binned": sd["rays_binned"]})
    ref = images[-1]
    for row, img in zip(rows, images):
        row["image_L1_vs_tightest"] = [float((i - r).abs().sum() / r.sum().clamp_min(1)) for i, r in zip(img, ref)]
        print(json.dumps(row))
    if a.out:
        with open(a.out, "w") as f:
            for row in rows:
                f.write(json.dumps(row) + "\n")


if __name__ == "__main__":
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device (no CPU fallback)")
    main()
