"""Ray bundle initial conditions with the call shape of ``src/simulator/beam.py::Beam``.

Small bundles are drawn on the host with NumPy's legacy global RNG in the reference's draw order, so a
seeded reference run and a seeded run here start from identical rays.  ``device=True`` keeps only the
parameters and lets the CUDA path generate ray i from (seed, i) with Philox -- required for >= 1e8 rays
(72 B/ray never touch HBM) and for partition-invariant multi-GPU runs.
"""
import numpy as np

from . import engine
from .engine import C_LIGHT as c


def random_array(length, seed=False):            # src/simulator/utils.py:8-12 (re-seeds before every draw)
    if seed:
        np.random.seed(0)
    return np.random.rand(length)


def random_array_n(length, seed=False):          # utils.py:14-18
    if seed:
        np.random.seed(0)
    return np.random.randn(length)


def random_inv_pow_array(power, length, seed=False):   # utils.py:20-24
    if seed:
        np.random.seed(0)
    return np.random.power(power, length)


class Beam:
    def __init__(self, Np, beam_size, divergence, ne_extent, *, probing_direction="z", wavelength=1064e-9,
                 beam_type="circular", seeded=False, device=False, seed=0):
        self.Np = int(Np)
        self.beam_size, self.divergence = beam_size, divergence
        self.probing_direction, self.beam_type, self.wavelength = probing_direction, beam_type, wavelength
        self.ne_extent = ne_extent
        self.device, self.seed = device, seed
        if device:
            self.s0 = None
            self.spec = engine.make_beam(beam_type, beam_size, divergence, ne_extent, probing_direction, seed)
        else:
            self.init_beam(ne_extent, seeded)

    def init_beam(self, ne_extent, seeded):
        """beam.py:35-303.  Draw order per type is the reference's."""
        Np, bs, div, pd, bt = self.Np, self.beam_size, self.divergence, self.probing_direction, self.beam_type
        s0 = np.zeros((9, Np))
        if bt == "circular":                              # beam.py:64-77
            t = 2 * np.pi * random_array(Np, seeded)
            u = random_array(Np, seeded)                  # drawn then overwritten upstream (beam.py:71-74)
            u = random_inv_pow_array(2, Np, seeded)
            phi = np.pi * random_array(Np)                # NOT seeded upstream (beam.py:76)
            chi = div * random_array_n(Np, seeded)
            a, b = bs * u * np.cos(t), bs * u * np.sin(t)
        elif bt == "even":
            # beam.py:210-227: concentric rings of 6 i points, i = 1..n_c, plus the centre.  Upstream builds the (u, t)
            # lists but never writes them into s0 (and its float ring count cannot drive range()), so there is nothing
            # to pin: here the rings are placed as that code intends (radius i/n_c, angle 2 pi j / (6 i)) and the
            # velocities come from the same chi / phi draws as the other types.  Np becomes 3 n_c (n_c + 1) + 1.
            n_c = int((-1 + np.sqrt(1 + 8 * (Np // 6))) / 2)
            Np = self.Np = 3 * (n_c + 1) * n_c + 1
            s0 = np.zeros((9, Np))
            phi = np.pi * random_array(Np, seeded)
            chi = div * random_array_n(Np, seeded)
            u, t = [0.0], [0.0]
            for i in range(1, n_c + 1):
                for j in range(6 * i):
                    u.append(i / n_c)
                    t.append(j * 2 * np.pi / (i * 6))
            u, t = np.array(u), np.array(t)
            a, b = bs * u * np.cos(t), bs * u * np.sin(t)
        elif bt in ("square", "rectangular", "rect_trackers"):   # beam.py:108-115, 150-162, 228-286 (rect_trackers == rectangular)
            t = 2 * random_array(Np, seeded) - 1.0
            u = 2 * random_array(Np, seeded) - 1.0
            phi = np.pi * random_array(Np, seeded)
            chi = div * random_array_n(Np, seeded)
            b1, b2 = (bs, bs) if bt == "square" else (bs[0], bs[1])
            a, b = b1 * u, b2 * t
        elif bt == "linear":                              # beam.py:195-208
            t = 2 * random_array(Np, seeded) - 1.0
            chi = div * random_array_n(Np, seeded)
            s0[3], s0[5] = c * np.sin(chi), c * np.cos(chi)
            s0[0], s0[2] = bs * t, -ne_extent
            s0[6] = 1.0
            self.s0 = s0
            return
        else:
            raise ValueError("beam_type unrecognised! Accepted args: circular, square, rectangular, rect_trackers, linear, even")
        para, p1, p2 = c * np.cos(chi), c * np.sin(chi) * np.cos(phi), c * np.sin(chi) * np.sin(phi)
        if pd == "x":
            s0[3], s0[4], s0[5] = para, p1, p2
            s0[0], s0[1], s0[2] = -ne_extent, a, b
        elif pd == "z":
            s0[3], s0[4], s0[5] = p1, p2, para
            s0[0], s0[1], s0[2] = a, b, -ne_extent
        else:
            s0[4], s0[3], s0[5] = para, p1, p2
            s0[0], s0[1], s0[2] = a, -ne_extent, b
        s0[6] = 1.0
        self.s0 = s0

    def materialise(self, n=None, ray_offset=0):
        """Device beams: generate rays [ray_offset, ray_offset+n) on the GPU as a (9,n) CUDA tensor."""
        if not self.device:
            raise RuntimeError("host beam: use .s0")
        return engine.beam_generate(self.spec, self.Np if n is None else n, ray_offset)

    def save_rays_pos(self, fn=None):                     # beam.py:305-321
        from datetime import datetime
        fn = "{} rays.npy".format(datetime.now().strftime("%Y-%m-%d_%H-%M-%S")) if fn is None else "{}.npy".format(fn)
        with open(fn, "wb") as f:
            np.save(f, self.s0)
