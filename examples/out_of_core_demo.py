#!/usr/bin/env python
"""Cost of tracing a grid slab by slab (synthpy_b200/out_of_core.py) against the in-core solve of the same grid: the C2
field of bench.py (turbulent 512^3 by default), device-generated rays, fixed-step RK4 at half a cell.  The two results
are compared bit for bit.  One JSON line on stdout.

    python examples/out_of_core_demo.py [--grid 512] [--rays 4e6] [--slab-planes 160] [--source device|host]

``--source host`` keeps the grid in host memory (what a real out-of-core run does: every slab is uploaded and packed);
``device`` slices a resident tensor, which isolates the cost of the extra launches and of re-packing.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from synthpy_b200 import beam as B, domain as Dm, out_of_core as OC, propagator as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--rays", type=float, default=4e6)
    ap.add_argument("--slab-planes", type=int, default=160)
    ap.add_argument("--source", choices=("device", "host"), default="host")
    a = ap.parse_args()
    args = bench.parse(["--workload", "C2", "--grid", str(a.grid)])
    ne = bench.build_ne(args, "cuda")
    dom = Dm.ScalarDomain(bench.LENGTHS, a.grid)
    dom.external_ne(ne)
    s0 = B.Beam(int(a.rays), bench.BEAM_R, bench.BEAM_DIV, bench.EXTENT, device=True, seed=2).materialise()
    rf, _, _, ex = P.solve(s0, dom, bench.EXTENT, lwl=bench.LWL, method="rk4", return_state=True)           # warm-up + reference
    rf, _, t_in, ex = P.solve(s0, dom, bench.EXTENT, lwl=bench.LWL, method="rk4", return_state=True)
    src = OC.array_source(ne if a.source == "device" else ne.cpu().numpy())
    out = None
    for _ in range(2):                                                                                       # warm-up, then timed
        out = OC.solve_out_of_core(s0, src, bench.LENGTHS, a.grid, bench.EXTENT, slab_planes=a.slab_planes, lwl=bench.LWL,
                                   return_state=True)
    rf2, _, t_oc, ex2 = out
    line = {"grid": a.grid, "rays": int(a.rays), "slab_planes": a.slab_planes, "source": a.source,
            "slabs": [dict(planes=list(e["planes"]), steps=e["steps"]) for e in ex2["slabs"]],
            "in_core_s": round(t_in, 4), "out_of_core_s": round(t_oc, 4), "ratio": round(t_oc / t_in, 3),
            "ray_steps": ex["stats"]["ray_steps"], "identical": bool(torch.equal(rf, rf2) and torch.equal(ex["sf"], ex2["sf"])
                                                                    and torch.equal(ex["steps"].long(), ex2["steps"].long()))}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
