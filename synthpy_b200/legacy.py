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


class MinimalScalarDomain:
    """``minimal_solver.ScalarDomain`` (src/solvers-legacy/minimal_solver.py:121-398), the 6-component generation of the
    legacy solver, on the CUDA path: state [x, y, z, vx, vy, vz], joint RK45 whose RMS error norm runs over 6N components
    (``sp_params.n_state = 6``), integration span sqrt(ex^2 + ey^2 ez^2) / c exactly as upstream writes it (:321), and the
    ``ne_max`` clamp of its ``calc_dndr`` (:231).

    Upstream keeps axes, density and gradients in float64 here; the device field is float32 ({g_u, g_v, g_w, aux} per
    node, the layout that matches ``full_solver`` bit for bit), so the float64 gradients are computed on the host with
    the reference's own NumPy call and rounded once when packed: results agree with ``minimal_solver`` to ~1e-7 relative
    (SURVEY.md 8a-3), not to 1e-9.  The step sequence (attempt count) is pinned by tests/golden/g8_minimal.npz."""

    def __init__(self, x, y, z, probing_direction="z"):
        self.x, self.y, self.z = (np.asarray(a, dtype=np.float64) for a in (x, y, z))
        self.extent_x, self.extent_y, self.extent_z = self.x.max(), self.y.max(), self.z.max()
        self.probing_direction = probing_direction
        self.extent = {"x": self.extent_x, "y": self.extent_y, "z": self.extent_z}[probing_direction]
        self.ne = None

    def _mesh(self):
        return np.meshgrid(self.x, self.y, self.z, indexing="ij")

    def test_null(self):                                            # minimal_solver.py:150-155
        self.ne = np.zeros((len(self.x), len(self.y), len(self.z)))

    def test_lens(self, n_e0=1e24, LR=1e-3):                        # minimal_solver.py:192-201
        XX, YY, _ = self._mesh()
        self.ne = n_e0 * np.exp(-(np.sqrt(XX ** 2 + YY ** 2)) ** 2 / LR ** 2)

    def test_liner(self, n_e0=1e24, LR=1e-3):                       # minimal_solver.py:203-212
        XX, _, ZZ = self._mesh()
        self.ne = n_e0 * np.exp(-(np.sqrt(XX ** 2 + ZZ ** 2)) ** 2 / LR ** 2)

    def external_ne(self, ne):                                      # minimal_solver.py:214-220
        self.ne = np.array(ne, dtype=np.float64, copy=True)

    def calc_dndr(self, lwl=1053e-9, ne_max=1):                     # minimal_solver.py:222-243
        self.omega = engine.omega_of(lwl)
        ne_nc = self.ne / engine.critical_density(self.omega)
        ne_nc[ne_nc > ne_max] = ne_max
        axes = (self.x, self.y, self.z)
        g = [-0.5 * engine.C_LIGHT ** 2 * np.gradient(ne_nc, axes[a], axis=a) for a in range(3)]
        self.field = engine.DeviceField.from_gradients(g[0], g[1], g[2], self.x, self.y, self.z,
                                                       march_axis=engine.AXIS[self.probing_direction])

    def init_beam(self, Np, beam_size, divergence):                 # minimal_solver.py:260-314 (legacy radial law, global RNG)
        self.s0 = init_beam(Np, beam_size, divergence, self.extent, "circular", self.probing_direction)[:6]
        return self.s0

    def t_end(self):
        return np.sqrt(self.extent_x ** 2 + self.extent_y ** 2 * self.extent_z ** 2) / engine.C_LIGHT

    def solve(self, method="RK45", s0=None):                        # minimal_solver.py:316-335
        s6 = np.asarray(self.s0 if s0 is None else s0, dtype=np.float64)
        s9 = np.zeros((9, s6.shape[1]))
        s9[:6] = s6
        P = engine.make_params("rk45_joint", probing_direction=self.probing_direction, extent=self.extent, omega=self.omega,
                               t_end=self.t_end(), rtol=1e-3, atol=1e-6, n_state=6, early_exit=False)
        out = engine.propagate(self.field, P, s0=engine.to_device(s9), want_sf=True, want_rf=True, want_steps=True)
        torch.cuda.synchronize()
        self.sf = out["sf"].cpu().numpy()[:6]
        self.rf = out["rf"].cpu().numpy()
        self.steps = out["steps"].cpu().numpy()
        self.stats = engine.stats_dict(out["stats_dev"])
        return self.rf

    def ray_at_exit(self):                                          # minimal_solver.py:337-384
        return self.rf


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
        super().coherent_solve(wl=wl, generation="legacy")


class Interferometry(_Rays, _diag.Interferometry):
    def two_lens_solve(self, wl=532e-9):                                                # rtm_solver.py:376
        super().two_lens_solve(wl=wl, ref_beam=None)
