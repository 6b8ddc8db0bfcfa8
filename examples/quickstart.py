#!/usr/bin/env python
"""The reference's canonical walkthrough (examples/notebooks/test_SynthRayTracer.ipynb, cells 2-15) on synthpy_b200:
same calls, same argument meaning -- only the import line changes.

    python examples/quickstart.py [n_rays]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from synthpy_b200 import beam as beam_initialiser, diagnostics as diag, domain as d, propagator as p  # noqa: E402


def main(Np=300000):
    # domain: 10 x 10 x 20 mm box, 128^3 cells, exponential-cos test profile (notebook cells 3-5)
    extent_x, extent_y, extent_z = 5e-3, 5e-3, 10e-3
    lengths = 2 * np.array([extent_x, extent_y, extent_z])
    probing_extent = extent_z
    domain = d.ScalarDomain(lengths, 128, ne_type="test_exponential_cos", probing_direction="z")

    # beam (cell 8)
    lwl = 1064e-9
    beam = beam_initialiser.Beam(Np, 5e-3, 5e-5, probing_extent, probing_direction="z", wavelength=lwl, beam_type="circular")

    # trace (cell 10)
    rf, Jf, duration = p.solve(beam.s0, domain, probing_extent, lwl=lwl)
    print(f"traced {Np} rays through 128^3 in {duration:.3f} s (reference notebook: 11.98 s on 16 CPU devices)")

    # diagnostics (cells 12-15)
    for name, cls, solve in (("refractometer", diag.Refractometry, "incoherent_solve"),
                             ("shadowgraphy", diag.Shadowgraphy, "single_lens_solve"),
                             ("schlieren", diag.Schlieren, "DF_solve")):
        t0 = time.time()
        o = cls(lwl, rf)
        getattr(o, solve)()
        o.histogram(bin_scale=1)
        print(f"{name:14s} H {o.H.shape}, {int(o.H.sum())} of {Np} rays on the detector  ({time.time() - t0:.3f} s)")
    return rf


if __name__ == "__main__":
    main(int(float(sys.argv[1])) if len(sys.argv) > 1 else 300000)
