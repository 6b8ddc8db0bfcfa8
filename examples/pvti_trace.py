#!/usr/bin/env python
"""The reference's production driver (examples/jobs/run_scripts/pvti_trace_multiprocess.py, pvti_trace_mpi.py) on
synthpy_b200: read an electron-density dump from a .pvti file, trace Np rays through it, write the shadowgraphy
and refractometer images.

    python examples/pvti_trace.py 1e8 field.pvti out_            # one GPU
    torchrun --nproc-per-node 8 examples/pvti_trace.py 1e9 field.pvti out_

What the reference does with 50 pool workers x 1e4-ray tasks and a Python `+=` over pickled histograms is here one
fused kernel per chunk of rays on each GPU (rays generated on the device, exit rays never stored) and one NCCL
all-reduce of the two images.  Outputs are pickled (ny, nx) float64 arrays under the same names.

    --demo N    writes a synthetic N^3 dump first (Gaussian column, in the driver's units) so the script runs as is
"""
import argparse
import os
import pickle
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from synthpy_b200 import beam as B, diagnostics as D, distributed, handle_filetypes as io, propagator as P  # noqa: E402


def write_demo(path, n):
    """A dump in the units the reference's drivers expect (ne in cm^-3 x 1e-6, lengths in m): Gaussian column."""
    ax = np.linspace(-1, 1, n)
    X, Y, _ = np.meshgrid(ax, ax, ax, indexing="ij")
    ne = 1e12 * np.exp(-(X ** 2 + Y ** 2) / 0.2 ** 2)
    io.export_pvti(ne.astype(np.float32), fname=path, extent_x=5e-3, extent_y=5e-3, extent_z=5e-3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("Np", type=float)
    ap.add_argument("file_loc")
    ap.add_argument("output_loc")
    ap.add_argument("--probing", default="y", choices=["x", "y", "z"])       # the reference's driver probes along y
    ap.add_argument("--scale", type=float, default=1e12)                      # field.external_ne(ne*1e12), driver line 58
    ap.add_argument("--divergence", type=float, default=0.05e-3)
    ap.add_argument("--chunk", type=float, default=5e7, help="rays per launch and GPU")
    ap.add_argument("--bin-scale", type=int, default=1)
    ap.add_argument("--demo", type=int, default=0)
    a = ap.parse_args()

    rank, world = distributed.init()
    if a.demo and rank == 0 and not os.path.exists(a.file_loc):
        write_demo(a.file_loc[:-5] if a.file_loc.endswith(".pvti") else a.file_loc, a.demo)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()

    t0 = time.time()
    dom, ext = io.domain_from_pvti(a.file_loc, probing_direction=a.probing, scale=a.scale, device="cuda")
    ax = {"x": 0, "y": 1, "z": 2}[a.probing]
    probing_extent = ext[ax]
    beam_size = ext[(ax + 1) % 3] if a.probing != "y" else ext[0]             # driver: beam_size = extent_x
    if rank == 0:
        print(f"extents: {ext}\ndims: {tuple(dom.dims)}\nfield loaded in {time.time() - t0:.2f} s")
    dom.device_field(1064e-9)
    dom.release_ne()

    Np = int(a.Np)
    beam = B.Beam(Np, beam_size, a.divergence, probing_extent, probing_direction=a.probing, beam_type="circular",
                  device=True, seed=0)
    specs = [D.spec("shadow_single", bin_scale=a.bin_scale), D.spec("refracto_incoherent", bin_scale=a.bin_scale)]
    off, cnt = distributed.shard(Np, rank, world)
    t0 = time.time()
    done = 0
    while done < cnt:                                                         # the reference's 1e4-ray tasks, 5e7 at a time
        n = min(int(a.chunk), cnt - done)
        stats, _ = P.solve_and_image(dom, beam, probing_extent, specs, n_rays=n, ray_offset=off + done, sync=False)
        done += n
    distributed.allreduce_images([s.image for s in specs])                    # comm.reduce(H, op=MPI.SUM)
    sh_H, r_H = (s.image.result().cpu().numpy() for s in specs)
    if rank == 0:
        dt = time.time() - t0
        print(f"{Np} rays on {world} GPU(s) in {dt:.2f} s ({Np / dt:.3g} rays/s); "
              f"{int(sh_H.sum())} / {int(r_H.sum())} rays on the shadowgraphy / refractometer detectors")
        with open(a.output_loc + "shadow.pkl", "wb") as fh:
            pickle.dump(sh_H, fh)
        with open(a.output_loc + "refract.pkl", "wb") as fh:
            pickle.dump(r_H, fh)
    return sh_H, r_H


if __name__ == "__main__":
    main()
