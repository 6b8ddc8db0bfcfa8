#!/usr/bin/env python
"""The reference's interferometry driver (examples/jobs/run_scripts/interference_MPI.py) on synthpy_b200: density dump
from a .pvti file, phase accumulated along every ray, reference beam added, two-lens telescope with E-field
propagation, interferogram; rays split over the ranks and the result summed on all of them.

    python examples/interference_mpi.py 1e7 1e6 field.pvti out_             # Np, scale_factor, file, output prefix
    torchrun --nproc-per-node 8 examples/interference_mpi.py 1e9 1e6 field.pvti out_

Two ways of adding chunks up:
  default            the complex amplitudes of ALL rays (every chunk, every rank) are summed per pixel and the magnitude
                     is taken once -- one coherent interferogram, independent of how the rays were split;
  --sum-magnitudes   what the reference's driver does (interference_MPI.py:160-189): every 1e6-ray chunk makes its own
                     interferogram H = sqrt(Re(sum Ex)^2 + Re(sum Ey)^2) and the magnitudes are added (`sh.H += ...`,
                     `comm.reduce(sh.H, op=MPI.SUM)`), so the image depends on the chunking.
"""
import argparse
import os
import pickle
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from synthpy_b200 import beam as B, diagnostics as D, distributed, handle_filetypes as io, propagator as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("Np", type=float)
    ap.add_argument("scale_factor", type=float)
    ap.add_argument("file_loc")
    ap.add_argument("output_loc")
    ap.add_argument("--probing", default="y", choices=["x", "y", "z"])
    ap.add_argument("--wl", type=float, default=532e-9)                         # driver line 69
    ap.add_argument("--beam-size", type=float, default=6e-3)
    ap.add_argument("--divergence", type=float, default=0.05e-3)
    ap.add_argument("--fringes", type=float, default=120.0)                     # s.interfere_ref_beam(rf, E, 120, -20)
    ap.add_argument("--deg", type=float, default=-20.0)
    ap.add_argument("--chunk", type=float, default=1e6, help="rays per chunk (Np_ray_split upstream)")
    ap.add_argument("--bin-scale", type=int, default=1)
    ap.add_argument("--sum-magnitudes", action="store_true")
    a = ap.parse_args()

    rank, world = distributed.init()
    dom, ext = io.domain_from_pvti(a.file_loc, probing_direction=a.probing, scale=a.scale_factor, device="cuda", phaseshift=True)
    ax = {"x": 0, "y": 1, "z": 2}[a.probing]
    depth = ext[ax]
    dom.device_field(a.wl, phase=True)
    dom.release_ne()

    Np = int(a.Np)
    beam = B.Beam(Np, a.beam_size, a.divergence, depth, probing_direction=a.probing, beam_type="circular", device=True, seed=0)
    spec = D.spec("interf_two", bin_scale=a.bin_scale, interferogram=True, wavelength=a.wl, ref_beam=(a.fringes, a.deg))
    off, cnt = distributed.shard(Np, rank, world)
    H = torch.zeros((spec.image.ny, spec.image.nx), dtype=torch.float64, device="cuda")
    t0 = time.time()
    done = 0
    while done < cnt:
        n = min(int(a.chunk), cnt - done)
        P.solve_and_image(dom, beam, depth, [spec], lwl=a.wl, n_rays=n, ray_offset=off + done, sync=False)
        done += n
        if a.sum_magnitudes:                       # per-chunk magnitude, as upstream
            H += spec.image.result()
            spec.image.zero_()
    if a.sum_magnitudes:
        distributed.allreduce_images([H])
    else:
        distributed.allreduce_images([spec.image])  # complex planes of all ranks, then one magnitude
        H = spec.image.result()
    H = H.cpu().numpy()
    if rank == 0:
        dt = time.time() - t0
        print(f"{Np} rays on {world} GPU(s) in {dt:.2f} s ({Np / dt:.3g} rays/s); interferogram {H.shape}, max {H.max():.4g}")
        with open(a.output_loc + "interferogram.pkl", "wb") as fh:
            pickle.dump(H, fh)
    return H


if __name__ == "__main__":
    main()
