#!/usr/bin/env python
"""Per-ray parity data on a benchmarked configuration: the CUDA path, the oracle, and the oracle on the same rays
perturbed by one ulp (its own conditioning), saved for offline analysis.
usage: parity_dump.py WORKLOAD N OFFSET out.npz"""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import bench
import torch
from oracle import parallel as OP, synthpy_oracle as O
from synthpy_b200 import beam as B, domain as Dm, propagator as P

w, n, off, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
a = bench.parse(["--workload", w])
ne = bench.build_ne(a, "cuda")
dom = Dm.ScalarDomain(bench.LENGTHS, a.grid)
dom.external_ne(ne)
odom = bench.cpu_setup(ne.cpu().numpy(), a, phaseshift=False)
del ne
beam = B.Beam(int(a.rays), bench.BEAM_R, bench.BEAM_DIV, bench.EXTENT, device=True, seed=2, beam_type="circular")
s0 = beam.materialise(n, off)
kw = bench.solve_kw(a, dom)
rf, _, _, ex = P.solve(s0, dom, bench.EXTENT, return_state=True, **kw)
h, ns = bench.rk4_lattice(a)
s0h = s0.cpu().numpy()
sf_o, st_o = OP.solve_rk4(odom, s0h, ns, h=h, early_exit=True)
s1 = s0h.copy()
s1[0] = np.nextafter(s1[0], np.inf); s1[1] = np.nextafter(s1[1], -np.inf)
sf_1, _ = OP.solve_rk4(odom, s1, ns, h=h, early_exit=True)
np.savez(out, rf=rf.cpu().numpy(), sf=ex["sf"].cpu().numpy(), steps=ex["steps"].cpu().numpy(), sf_o=sf_o, steps_o=st_o, sf_1=sf_1,
         rf_o=O.ray_to_jones(sf_o, bench.EXTENT)[0], rf_1=O.ray_to_jones(sf_1, bench.EXTENT)[0], s0=s0h)
r = np.abs(rf.cpu().numpy() - O.ray_to_jones(sf_o, bench.EXTENT)[0])
print("saved", out, "max abs diff rows", r.max(axis=1))
