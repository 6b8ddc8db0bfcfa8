"""Legacy-shaped aliases (``src/solvers-legacy/full_solver.py`` + ``rtm_solver.py`` call shapes) over the same
CUDA path.  The parity tests use these so that they read like the reference's own calls:

    dom = legacy.ScalarDomain(x, y, z, extent); dom.external_ne(ne); dom.calc_dndr(lwl)
    rf = dom.solve(s0)                        # the reference as shipped: joint RK45
    sh = legacy.Shadowgraphy(rf, L=400, R=25); sh.single_lens_solve(); sh.histogram(bin_scale=10)
"""
import numpy as np
import torch

from . import diagnostics as _diag
from . import engine
from .engine import C_LIGHT as c


class ScalarDomain:
    def __init__(self, x, y, z, extent, B_on=False, inv_brems=False, phaseshift=False, probing_direction="z"):
        self.B_on, self.inv_brems = B_on, inv_brems
        self.B = self.Te = self.Z = None
        self.x, self.y, self.z = np.float32(x), np.float32(y), np.float32(z)       # full_solver.py:119
        self._axes64 = (np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), np.asarray(z, dtype=np.float64))
        self.extent, self.probing_direction, self.phaseshift = extent, probing_direction, phaseshift
        self.ne = None
        self.field = None

    def _mesh(self):
        return np.meshgrid(*self._axes64, indexing="ij", copy=False)

    def test_null(self):
        self.ne = np.zeros_like(self._mesh()[0])

    def test_slab(self, s=1, n_e0=2e23):
        self.ne = n_e0 * (1.0 + s * self._mesh()[0] / self.extent)

    def test_linear_cos(self, s1=0.1, s2=0.1, n_e0=2e23, Ly=1):
        XX, YY, _ = self._mesh()
        self.ne = n_e0 * (1.0 + s1 * XX / self.extent) * (1 + s2 * np.cos(2 * np.pi * YY / Ly))

    def test_exponential_cos(self, n_e0=1e24, Ly=1e-3, s=2e-3):
        XX, YY, _ = self._mesh()
        self.ne = n_e0 * 10 ** (XX / s) * (1 + np.cos(2 * np.pi * YY / Ly))

    def test_lens(self, n_e0=1e24, LR=1e-3):        # minimal_solver.py:192-201
        XX, YY, _ = self._mesh()
        self.ne = n_e0 * np.exp(-(np.sqrt(XX ** 2 + YY ** 2)) ** 2 / LR ** 2)

    def test_liner(self, n_e0=1e24, LR=1e-3):       # minimal_solver.py:203-212
        XX, _, ZZ = self._mesh()
        self.ne = n_e0 * np.exp(-(np.sqrt(XX ** 2 + ZZ ** 2)) ** 2 / LR ** 2)

    def external_ne(self, ne):
        self.ne = ne

    def external_B(self, B):                      # full_solver.py:177-183
        self.B = B

    def external_Te(self, Te, Te_min=1.0):        # full_solver.py:185-191
        self.Te = np.maximum(Te_min, Te)

    def external_Z(self, Z):                      # full_solver.py:193-199
        self.Z = Z

    def test_B(self, Bmax=1.0):                   # full_solver.py:201-209
        XX = self._mesh()[0]
        self.B = np.zeros(XX.shape + (3,))
        self.B[..., 2] = Bmax * XX / self.extent

    def set_up_interps(self):
        """full_solver.py:276-289: attach the attenuation / Faraday grids to the device field."""
        kappa = engine.kappa_grid(self.ne, self.Te, self.Z, self.omega) if self.inv_brems else None
        self.field.attach_channels(kappa=kappa, ne=self.ne if self.B_on else None, B=self.B if self.B_on else None)

    def calc_dndr(self, lwl=1053e-9, phase_f64=True, ne_max=None):
        """full_solver.py:211-234 on the device (float32 stencil identical to np.gradient).  ``ne_max`` (in units of
        the critical density) is minimal_solver.calc_dndr's clamp, minimal_solver.py:231: ne_nc[ne_nc > ne_max] = ne_max."""
        self.lwl = lwl
        self.omega = engine.omega_of(lwl)
        self.VerdetConst = 2.62e-13 * lwl ** 2 if self.B_on else 0.0          # full_solver.py:222-223
        ne = self.ne
        if ne_max is not None:
            cap = ne_max * engine.critical_density(self.omega)
            ne = ne.clamp(max=cap) if isinstance(ne, torch.Tensor) else np.minimum(ne, cap)
        self.field = engine.DeviceField.from_ne(ne, self.x, self.y, self.z, self.omega,
                                                march_axis=engine.AXIS[self.probing_direction],
                                                phase=self.phaseshift, phase_f64=self.phaseshift and phase_f64)

    def params(self, method, **kw):
        kw.setdefault("phase", self.phaseshift)
        kw.setdefault("phase_f64", self.phaseshift and self.field.has_f64)
        kw.setdefault("atten", self.inv_brems)
        kw.setdefault("faraday", self.B_on)
        kw.setdefault("verdet", self.VerdetConst)
        return engine.make_params(method, probing_direction=self.probing_direction, extent=self.extent,
                                  omega=self.omega, **kw)

    def dsdt(self, s):
        """fs.dsdt(t, s, dom) for a (9,N) state; returns (9,N) numpy."""
        sd = engine.to_device(np.asarray(s).reshape(9, -1))
        return engine.rhs(self.field, self.params("rk4", n_steps=1, h=1.0), sd).cpu().numpy()

    def solve(self, s0, return_E=False, method="rk45_joint", rtol=1e-3, atol=1e-6, n_steps=0, h=0.0,
              early_exit=False, fp32=False, sort=True):
        """full_solver.py:376-403.  Default = the shipped algorithm (one step size for all rays)."""
        P = self.params(method, rtol=rtol, atol=atol, n_steps=n_steps, h=h, early_exit=early_exit, fp32=fp32,
                        sort=sort)
        out = engine.propagate(self.field, P, s0=engine.to_device(s0), want_sf=True, want_rf=True, want_jf=True,
                               want_steps=True)
        torch.cuda.synchronize()
        self.sf = out["sf"].cpu().numpy()
        self.rf = out["rf"].cpu().numpy()
        self.Jf = out["jf"].cpu().numpy()
        self.steps = out["steps"].cpu().numpy()
        self.stats = engine.stats_dict(out["stats_dev"])
        return (self.rf, self.Jf) if return_E else self.rf


def init_beam(Np, beam_size, divergence, ne_extent, beam_type="circular", probing_direction="z"):
    """full_solver.py:547-835 on the host RNG (legacy radial law u = fold(U+U)); same draw order."""
    s0 = np.zeros((9, Np))
    R = np.random
    if beam_type == "circular":
        t = 2 * np.pi * R.rand(Np)
        u = R.rand(Np) + R.rand(Np)
        u[u > 1] = 2 - u[u > 1]
        phi, chi = np.pi * R.rand(Np), divergence * R.randn(Np)
        a, b = beam_size * u * np.cos(t), beam_size * u * np.sin(t)
    elif beam_type in ("square", "rectangular"):
        t, u = 2 * R.rand(Np) - 1.0, 2 * R.rand(Np) - 1.0
        phi, chi = np.pi * R.rand(Np), divergence * R.randn(Np)
        b1, b2 = (beam_size, beam_size) if beam_type == "square" else (beam_size[0], beam_size[1])
        a, b = b1 * u, b2 * t
    elif beam_type == "linear":
        t = 2 * R.rand(Np) - 1.0
        chi = divergence * R.randn(Np)
        s0[3], s0[5], s0[0], s0[2], s0[6] = c * np.sin(chi), c * np.cos(chi), beam_size * t, -ne_extent, 1.0
        return s0
    else:
        raise ValueError("beam_type unrecognised")
    para, p1, p2 = c * np.cos(chi), c * np.sin(chi) * np.cos(phi), c * np.sin(chi) * np.sin(phi)
    if probing_direction == "x":
        s0[3], s0[4], s0[5], s0[0], s0[1], s0[2] = para, p1, p2, -ne_extent, a, b
    elif probing_direction == "z":
        s0[3], s0[4], s0[5], s0[0], s0[1], s0[2] = p1, p2, para, a, b, -ne_extent
    else:
        s0[4], s0[3], s0[5], s0[0], s0[1], s0[2] = para, p1, p2, a, -ne_extent, b
    s0[6] = 1.0
    return s0


# ---- rtm_solver.py call shapes ------------------------------------------------------------------------------
class _Rays:
    def __init__(self, r0, E=None, focal_plane=0, L=400, R=25, Lx=18, Ly=13.5):       # rtm_solver.py:142-153
        super().__init__(None, r0, E, focal_plane=focal_plane, L=L, R=R, Lx=Lx, Ly=Ly)

    def histogram(self, bin_scale=10, pix_x=3448, pix_y=2574, clear_mem=False):         # rtm_solver.py:156
        super().histogram(bin_scale=bin_scale, pix_x=pix_x, pix_y=pix_y, clear_mem=clear_mem)

    @property
    def rE(self):
        return self.Jf


class Shadowgraphy(_Rays, _diag.Shadowgraphy):
    pass


class Schlieren(_Rays, _diag.Schlieren):
    pass


class Refractometry(_Rays, _diag.Refractometry):
    def coherent_solve(self, wl=1064e-9):
        super().coherent_solve(wl=wl)


class Interferometry(_Rays, _diag.Interferometry):
    def two_lens_solve(self, wl=532e-9):                                                # rtm_solver.py:376
        super().two_lens_solve(wl=wl, ref_beam=None)
