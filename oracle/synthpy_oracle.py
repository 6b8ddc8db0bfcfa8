"""CPU oracle for the synthPy ray-propagation hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a NumPy/SciPy restatement of the reference's *legacy* (runnable) generation of the hot
path: field preparation -> ray ODE -> exit-plane projection -> ray-transfer-matrix optics -> detector
binning.  It exists to check the CUDA path; nothing under ``synthpy_b200/`` may import it.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs use it.

Parity status: **pinned**.  ``oracle/gen_golden.py`` imports the real reference
(``/root/reference/src/solvers-legacy/{full_solver,rtm_solver}.py``) in the build container, runs it on
seeded inputs and commits the outputs under ``tests/golden/``;  ``tests/test_oracle_golden.py`` holds this
restatement to those vectors (bit-exact for the RHS / joint RK45 / optics / histograms).  The reference has
no test-suite of its own; its two docstring known-answer cases (NULL and SLAB test,
full_solver.py:12-82) are covered in the same test file.  The optics / detector code of the current generation
(src/simulator/diagnostics.py) is pinned the same way (g9: its source run with NumPy standing in for jax.numpy);
the one restatement without a fixture, the diffrax Tsit5 + PID solve, says PARITY UNPINNED where it is defined.

Third-party arithmetic the reference calls and which is therefore called here at the same call sites:
``scipy.interpolate.RegularGridInterpolator`` (full_solver.py:232-234,289,344) and
``scipy.integrate.solve_ivp`` RK45 (full_solver.py:391).  Pinned by the reference at scipy==1.13.1 /
numpy==1.26.4 (MAGPIE_venv.yml:234,96); this image has scipy 1.18.1 / numpy 2.3.5 (same algorithm constants).

All ``file:line`` citations are relative to /root/reference/src/solvers-legacy/ unless stated otherwise.
"""
from __future__ import annotations

import numpy as np
from scipy.constants import c as C_LIGHT
from scipy.integrate import solve_ivp
from scipy.interpolate import RegularGridInterpolator

# --------------------------------------------------------------------------------------------------
# Field container + RHS                                                         (full_solver.py:96-544)
# --------------------------------------------------------------------------------------------------

NC_COEFF = 3.14207787e-4      # full_solver.py:219   n_c = NC_COEFF * omega^2
OMEGA_PE_COEFF = 5.64e4       # full_solver.py:239   omega_pe = 5.64e4 sqrt(ne[cm^-3])


class Domain:
    """Restates ``full_solver.ScalarDomain`` (full_solver.py:96-403): float32 axes, float32 normalised
    density and gradients, float64 interpolation arithmetic, 9-component ray state."""

    def __init__(self, x, y, z, extent, *, phaseshift=False, probing_direction="z", B_on=False, inv_brems=False):
        # full_solver.py:119 -- the axes are *rounded to float32*; the mesh (used by analytic profiles
        # only) is built from the caller's float64 axes (full_solver.py:120).
        self.x, self.y, self.z = (np.float32(a) for a in (x, y, z))
        self._mesh_axes = (np.asarray(x), np.asarray(y), np.asarray(z))
        self.extent = extent
        self.probing_direction = probing_direction
        self.phaseshift = phaseshift
        self.B_on, self.inv_brems = B_on, inv_brems
        self.ne = self.B = self.Te = self.Z = None

    # -- analytic profiles (full_solver.py:130-175) -------------------------------------------------
    def _mesh(self):
        return np.meshgrid(*self._mesh_axes, indexing="ij", copy=False)

    def test_null(self):                                   # full_solver.py:130-134
        self.ne = np.zeros_like(self._mesh()[0])

    def test_slab(self, s=1, n_e0=2e23):                   # full_solver.py:136-146
        self.ne = n_e0 * (1.0 + s * self._mesh()[0] / self.extent)

    def test_linear_cos(self, s1=0.1, s2=0.1, n_e0=2e23, Ly=1):   # full_solver.py:148-157
        XX, YY, _ = self._mesh()
        self.ne = n_e0 * (1.0 + s1 * XX / self.extent) * (1 + s2 * np.cos(2 * np.pi * YY / Ly))

    def test_exponential_cos(self, n_e0=1e24, Ly=1e-3, s=2e-3):   # full_solver.py:159-167
        XX, YY, _ = self._mesh()
        self.ne = n_e0 * 10 ** (XX / s) * (1 + np.cos(2 * np.pi * YY / Ly))

    def external_ne(self, ne):                             # full_solver.py:169-175
        self.ne = ne

    def external_B(self, B):                               # full_solver.py:177-183
        self.B = B

    def external_Te(self, Te, Te_min=1.0):                 # full_solver.py:185-191
        self.Te = np.maximum(Te_min, Te)

    def external_Z(self, Z):                               # full_solver.py:193-199
        self.Z = Z

    def kappa(self):
        """Inverse-bremsstrahlung rate grid, full_solver.py:243-268 (NRL formulary)."""
        e = 1.602176634e-19                                # scipy.constants.e
        ne_cc = self.ne * 1e-6
        o_pe = OMEGA_PE_COEFF * np.sqrt(ne_cc)
        o_max = np.copy(o_pe)
        o_max[o_pe < self.omega] = self.omega
        L_max = np.maximum(self.Z * e / self.Te, 2.760428269727312e-10 / np.sqrt(self.Te))
        CL = np.maximum(2.0, np.log(4.19e5 * np.sqrt(self.Te) / (o_max * L_max)))
        return 3.1e-5 * self.Z * C_LIGHT * np.power(ne_cc / self.omega, 2) * CL * np.power(self.Te, -1.5)

    def set_up_interps(self):
        """full_solver.py:276-289 (the pieces the ray ODE uses)."""
        axes = (self.x, self.y, self.z)
        if self.B_on:
            self.verdet = 2.62e-13 * self.lwl ** 2         # full_solver.py:222-223
            self.ne_interp = RegularGridInterpolator(axes, self.ne, bounds_error=False, fill_value=0.0)
            self.B_interp = [RegularGridInterpolator(axes, self.B[..., c], bounds_error=False, fill_value=0.0) for c in range(3)]
        if self.inv_brems:
            self.kappa_interp = RegularGridInterpolator(axes, self.kappa(), bounds_error=False, fill_value=0.0)

    # -- gradient precompute (full_solver.py:211-234) -----------------------------------------------
    def calc_dndr(self, lwl=1053e-9):
        self.lwl = lwl
        self.omega = 2 * np.pi * (C_LIGHT / lwl)
        nc = NC_COEFF * self.omega ** 2
        self.ne_nc = np.array(self.ne / nc, dtype=np.float32)
        axes = (self.x, self.y, self.z)
        # float32 field, float32 axes -> np.gradient stays in float32 and takes the non-uniform branch
        self.grads = [-0.5 * C_LIGHT ** 2 * np.gradient(self.ne_nc, axes[a], axis=a) for a in range(3)]
        self.grad_interp = [
            RegularGridInterpolator(axes, g, bounds_error=False, fill_value=0.0) for g in self.grads
        ]
        if self.phaseshift:
            # full_solver.py:270-274,289,344: n = sqrt(1 - (omega_pe(ne*1e-6)/omega)^2), fill value 1.0.
            # (the reference rebuilds this interpolator on every RHS call; building it once is identical)
            n = np.sqrt(1.0 - (OMEGA_PE_COEFF * np.sqrt(self.ne * 1e-6) / self.omega) ** 2)
            self.n_interp = RegularGridInterpolator(axes, n, bounds_error=False, fill_value=1.0)

    # -- RHS (full_solver.py:317-347, 516-544) --------------------------------------------------------
    def dndr(self, pos):
        """pos: (3,N) -> (3,N) acceleration, full_solver.py:317-332."""
        pts = pos.T
        return np.stack([f(pts) for f in self.grad_interp])

    def dsdt(self, t, s):
        """Flattened 9N -> 9N, full_solver.py:516-544."""
        n = s.size // 9
        s = s.reshape(9, n)
        out = np.zeros_like(s)
        out[3:6] = self.dndr(s[:3])
        out[:3] = s[3:6]
        if self.inv_brems:
            out[6] = self.kappa_interp(s[:3].T) * s[6]                # full_solver.py:335-339,540
        if self.phaseshift:
            out[7] = self.omega * (self.n_interp(s[:3].T) - 1.0)      # full_solver.py:342-345
        if self.B_on:                                                 # full_solver.py:356-374,542
            Bv = np.sum(np.array([f(s[:3].T) for f in self.B_interp]) * s[3:6], axis=0)
            out[8] = self.verdet * self.ne_interp(s[:3].T) * Bv
        return out.ravel()

    # -- integrators ----------------------------------------------------------------------------------
    def t_end(self):
        return np.sqrt(8.0) * self.extent / C_LIGHT          # full_solver.py:381

    def solve_joint(self, s0, rtol=1e-3, atol=1e-6, return_stats=False):
        """The reference's shipped solve: ONE adaptive RK45 over the flattened 9N state
        (full_solver.py:376-403).  Returns the 9xN final state (``self.sf`` upstream)."""
        t = np.linspace(0.0, self.t_end(), 2)
        sol = solve_ivp(self.dsdt, [0, t[-1]], s0.ravel(), t_eval=t, rtol=rtol, atol=atol)
        sf = sol.y[:, -1].reshape(9, s0.shape[1])
        return (sf, sol) if return_stats else sf

    def solve_per_ray(self, s0, rtol=1e-3, atol=1e-6):
        """Same solver, one ray at a time (== ``ScalarDomain.solve`` called with Np=1 for each ray);
        the per-ray adaptive CUDA mode is held to this.  Returns (9xN state, nfev per ray)."""
        n = s0.shape[1]
        sf = np.empty((9, n))
        nfev = np.empty(n, dtype=np.int64)
        t = np.linspace(0.0, self.t_end(), 2)
        for i in range(n):
            sol = solve_ivp(self.dsdt, [0, t[-1]], s0[:, i].copy(), t_eval=t, rtol=rtol, atol=atol)
            sf[:, i] = sol.y[:, -1]
            nfev[i] = sol.nfev
        return sf, nfev

    def solve_rk4(self, s0, n_steps, h=None, early_exit=False):
        """Classical fixed-step RK4 whose RHS *is* the reference RHS (SURVEY 7.2 level L1).  The
        reference ships no fixed-step integrator; t-span and post-processing follow full_solver.py:376-403.
        ``early_exit`` freezes a ray once it is outside the grid on some axis and moving away from it
        (its RHS is identically zero from then on, so the frozen ray is on the same straight line).
        Returns (9xN state, steps taken per ray)."""
        if h is None:
            h = self.t_end() / n_steps
        s = np.array(s0, dtype=np.float64, copy=True)
        n = s.shape[1]
        live = np.ones(n, dtype=bool)
        steps = np.zeros(n, dtype=np.int64)
        lo = np.array([a[0] for a in (self.x, self.y, self.z)], dtype=np.float64)[:, None]
        hi = np.array([a[-1] for a in (self.x, self.y, self.z)], dtype=np.float64)[:, None]
        for _ in range(n_steps):
            if early_exit:
                p, v = s[:3, live], s[3:6, live]
                gone = np.any(((p > hi) & (v >= 0)) | ((p < lo) & (v <= 0)), axis=0)
                idx = np.flatnonzero(live)
                live[idx[gone]] = False
                if not live.any():
                    break
            y = s[:, live].ravel()
            k1 = self.dsdt(0.0, y)
            k2 = self.dsdt(0.0, y + (0.5 * h) * k1)
            k3 = self.dsdt(0.0, y + (0.5 * h) * k2)
            k4 = self.dsdt(0.0, y + h * k3)
            y = y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
            s[:, live] = y.reshape(9, -1)
            steps[live] += 1
        return s, steps


# --------------------------------------------------------------------------------------------------
# Tsit5 + PID controller (the current generation's solver)              PARITY UNPINNED: see below
# --------------------------------------------------------------------------------------------------
TSIT5_C = np.array([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
TSIT5_A = np.zeros((7, 7))
TSIT5_A[1, :1] = [0.161]
TSIT5_A[2, :2] = [-0.008480655492356989, 0.335480655492357]
TSIT5_A[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
TSIT5_A[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
TSIT5_A[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]
TSIT5_A[6, :6] = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]
TSIT5_B = TSIT5_A[6].copy()
TSIT5_BT = np.array([-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
                     0.5823571654525552, -0.45808210592918697, 0.015151515151515152])


def solve_tsit5_per_ray(dom, s0, rtol=1.0, atol=1e-5, save_steps=2, max_steps=10000):
    """``src/simulator/propagator.py:533-599``: every ray on its own with diffrax ``Tsit5`` under
    ``PIDController(rtol, atol)`` in normalised time tau = t / T, T = sqrt(8) extent / c, ``dt0 = T / save_steps``
    (tau units, as upstream writes ``(t1 - t0) * norm_factor / Nt``), ``max_steps = 10000``.

    PARITY UNPINNED: jax / diffrax / equinox are not installable in the build container, so this is a restatement of
    the published algorithm, not a recording of the library: Tsitouras' 5(4) pair (the tableau above passes the order
    conditions, tests/test_host_misc.py) in plain 9-vector form around the reference RHS ``Domain.dsdt``, and the
    controller diffrax documents for its defaults (pcoeff 0, icoeff 1, dcoeff 0; safety 0.9; factor limits 0.2 / 10,
    the lower one raised to 1 after an accepted step; rms norm of err / (atol + rtol max(|y0|, |y1|))).
    Returns (9xN state, attempted steps per ray, accepted steps per ray)."""
    T = dom.t_end()
    n = s0.shape[1]
    sf = np.empty((9, n)); att = np.zeros(n, dtype=np.int64); acc = np.zeros(n, dtype=np.int64)
    for i in range(n):
        y = np.array(s0[:, i], dtype=np.float64)
        f = dom.dsdt(0.0, y.copy())
        tau, dt = 0.0, T / save_steps
        while tau < 1.0 and att[i] < max_steps:
            last = tau + dt >= 1.0
            d = 1.0 - tau if last else dt
            h = d * T
            K = np.zeros((7, 9)); K[0] = f
            for s_ in range(1, 6):
                K[s_] = dom.dsdt(0.0, y + h * (TSIT5_A[s_, :s_] @ K[:s_]))
            yn = y + h * (TSIT5_B[:6] @ K[:6])
            K[6] = dom.dsdt(0.0, yn)
            err = h * (TSIT5_BT @ K)
            en = np.sqrt(np.mean((err / (atol + rtol * np.maximum(np.abs(y), np.abs(yn)))) ** 2))
            att[i] += 1
            keep = en < 1.0
            fac = 10.0 if en == 0 else min(10.0, max(1.0 if keep else 0.2, 0.9 * en ** -0.2))
            dt = d * fac
            if keep:
                tau = 1.0 if last else tau + d
                y, f = yn, K[6]
                acc[i] += 1
        sf[:, i] = y
    return sf, att, acc


class MinimalDomain:
    """Restates ``minimal_solver.ScalarDomain`` (src/solvers-legacy/minimal_solver.py:121-398), the 6-component
    generation of the solver: float64 axes, float64 normalised density clamped at ``ne_max`` critical densities
    (:231), float64 gradients, state [x, y, z, vx, vy, vz], RMS error norm over 6N components, and its own
    integration span sqrt(ex^2 + ey^2 ez^2) / c (:321, as written upstream)."""

    def __init__(self, x, y, z, probing_direction="z"):
        self.x, self.y, self.z = (np.asarray(a, dtype=np.float64) for a in (x, y, z))
        self.extent_x, self.extent_y, self.extent_z = self.x.max(), self.y.max(), self.z.max()
        self.probing_direction = probing_direction
        self.extent = {"x": self.extent_x, "y": self.extent_y, "z": self.extent_z}[probing_direction]

    def external_ne(self, ne):
        self.ne = np.array(ne, dtype=np.float64, copy=True)

    def calc_dndr(self, lwl=1053e-9, ne_max=1):                      # minimal_solver.py:222-243
        self.omega = 2 * np.pi * (C_LIGHT / lwl)
        nc = NC_COEFF * self.omega ** 2
        ne_nc = self.ne / nc
        ne_nc[ne_nc > ne_max] = ne_max
        axes = (self.x, self.y, self.z)
        self.grads = [-0.5 * C_LIGHT ** 2 * np.gradient(ne_nc, axes[a], axis=a) for a in range(3)]
        self.grad_interp = [RegularGridInterpolator(axes, g, bounds_error=False, fill_value=0.0) for g in self.grads]

    def dsdt(self, t, s):                                           # minimal_solver.py:506-527
        s = s.reshape(6, -1)
        out = np.zeros_like(s)
        out[3:6] = np.stack([f(s[:3].T) for f in self.grad_interp])
        out[:3] = s[3:6]
        return out.ravel()

    def t_end(self):
        return np.sqrt(self.extent_x ** 2 + self.extent_y ** 2 * self.extent_z ** 2) / C_LIGHT      # minimal_solver.py:321

    def solve(self, s0):                                            # minimal_solver.py:316-335
        t = np.linspace(0.0, self.t_end(), 2)
        sol = solve_ivp(self.dsdt, [0, t[-1]], np.asarray(s0, dtype=np.float64).ravel(), t_eval=t, method="RK45")
        self.sf, self.nfev = sol.y[:, -1].reshape(6, -1), sol.nfev
        return self.ray_at_exit()

    def ray_at_exit(self):                                          # minimal_solver.py:337-384
        s9 = np.zeros((9, self.sf.shape[1]))
        s9[:6] = self.sf
        return ray_to_jones(s9, self.extent, self.probing_direction)[0]


def init_beam(Np, beam_size, divergence, ne_extent, beam_type="circular", probing_direction="z", rng=None):
    """Restates ``full_solver.init_beam`` (full_solver.py:547-835): legacy radial law u = fold(U+U).
    Draw order (t, u1, u2, phi, chi) is the reference's so ``np.random.seed(k)`` reproduces its rays;
    pass ``rng=np.random`` (default) for the legacy global stream."""
    R = np.random if rng is None else rng
    s0 = np.zeros((9, Np))
    if beam_type == "circular":                                         # full_solver.py:565-572
        t = 2 * np.pi * R.rand(Np)
        u = R.rand(Np) + R.rand(Np)
        u[u > 1] = 2 - u[u > 1]
        phi = np.pi * R.rand(Np)
        chi = divergence * R.randn(Np)
        a, b = beam_size * u * np.cos(t), beam_size * u * np.sin(t)
    elif beam_type in ("square", "rectangular"):                        # full_solver.py:612-618,658-667
        t = 2 * R.rand(Np) - 1.0
        u = 2 * R.rand(Np) - 1.0
        phi = np.pi * R.rand(Np)
        chi = divergence * R.randn(Np)
        b1, b2 = (beam_size, beam_size) if beam_type == "square" else (beam_size[0], beam_size[1])
        a, b = b1 * u, b2 * t
    elif beam_type == "linear":                                         # full_solver.py:707-720
        t = 2 * R.rand(Np) - 1.0
        chi = divergence * R.randn(Np)
        s0[3], s0[4], s0[5] = C_LIGHT * np.sin(chi), 0.0, C_LIGHT * np.cos(chi)
        s0[0], s0[1], s0[2] = beam_size * t, 0.0, -ne_extent
        s0[6] = 1.0
        return s0
    else:
        raise ValueError("beam_type unrecognised")
    para = C_LIGHT * np.cos(chi)
    p1 = C_LIGHT * np.sin(chi) * np.cos(phi)
    p2 = C_LIGHT * np.sin(chi) * np.sin(phi)
    if probing_direction == "x":                                        # full_solver.py:574-582
        s0[3], s0[4], s0[5] = para, p1, p2
        s0[0], s0[1], s0[2] = -ne_extent, a, b
    elif probing_direction == "z":                                      # full_solver.py:592-600
        s0[3], s0[4], s0[5] = p1, p2, para
        s0[0], s0[1], s0[2] = a, b, -ne_extent
    else:                                                               # 'y' and the fall-through default
        s0[4], s0[3], s0[5] = para, p1, p2
        s0[0], s0[1], s0[2] = a, -ne_extent, b
    s0[6] = 1.0                                                         # full_solver.py:801-802,834
    return s0


def ray_to_jones(sf, ne_extent, probing_direction="z"):
    """Restates ``full_solver.ray_to_Jonesvector`` (full_solver.py:838-894): back-project to the exit
    plane, small-angle free arctan angles, Jones vector from (amp, phase, pol)."""
    ax = {"x": (0, 1, 2), "y": (1, 0, 2), "z": (2, 0, 1)}[probing_direction]
    p, a, b = ax
    t_bp = (sf[p] - ne_extent) / sf[3 + p]
    rp = np.zeros((4, sf.shape[1]))
    rp[0] = sf[a] - sf[3 + a] * t_bp
    rp[2] = sf[b] - sf[3 + b] * t_bp
    rp[1] = np.arctan(sf[3 + a] / sf[3 + p])
    rp[3] = np.arctan(sf[3 + b] / sf[3 + p])
    amp, phase, pol = sf[6], sf[7], sf[8]
    rot = amp * (np.cos(phase) + 1.0j * np.sin(phase))
    rJ = np.zeros((2, sf.shape[1]), dtype=complex)
    rJ[0] = rot * (-np.sin(pol))             # E_x_init = 0, E_y_init = 1 (full_solver.py:886-890)
    rJ[1] = rot * np.cos(pol)
    return rp, rJ


# --------------------------------------------------------------------------------------------------
# Ray-transfer-matrix optics                                                      (rtm_solver.py:48-453)
# --------------------------------------------------------------------------------------------------
# An optical train is a list of ops; the same list drives the CUDA epilogue, so tests read 1:1.
#   ("travel", d) | ("travel_noE", d) | ("lens", f1, f2) | ("circ_ap", R) | ("circ_stop", R) | ("rect_ap", Lx, Ly)
#   | ("knife", offset, axis 0|2, direction +-1)

def m_to_mm(r):                                                          # rtm_solver.py:48-51
    rr = np.array(r, copy=True)
    rr[0::2] *= 1e3
    return rr


def _blockdiag(m1, m2):
    L = np.zeros((4, 4))
    L[:2, :2], L[2:, 2:] = m1, m2
    return L


def apply_op(r, op):
    """One optical element on a (4,N) ray bundle; rejected rays become NaN columns.  Matrices are
    applied as ``np.matmul(L, r)`` exactly as the reference does (rtm_solver.py:53-136)."""
    kind = op[0]
    if kind in ("travel", "travel_noE"):                                 # rtm_solver.py:73-82
        d = np.array([[1, op[1]], [0, 1]])
        return np.matmul(_blockdiag(d, d), r)
    if kind == "lens":                                                   # rtm_solver.py:53-65
        l1 = np.array([[1, 0], [-1 / op[1], 1]])
        l2 = np.array([[1, 0], [-1 / op[2], 1]])
        return np.matmul(_blockdiag(l1, l2), r)
    r = np.array(r, copy=True)
    rr = r[0] ** 2 + r[2] ** 2
    if kind == "circ_ap":                                                # rtm_solver.py:84-90
        filt = rr > op[1] ** 2
    elif kind == "circ_stop":                                            # rtm_solver.py:92-98
        filt = rr < op[1] ** 2
    elif kind == "rect_ap":                                              # rtm_solver.py:110-118 (AND quirk)
        filt = (r[0] ** 2 > op[1] ** 2) * (r[2] ** 2 > op[2] ** 2)
    elif kind == "knife":                                                # rtm_solver.py:120-136
        filt = r[op[2]] > op[1] if op[3] > 0 else r[op[2]] < op[1]
    else:
        raise ValueError(kind)
    r[:, filt] = np.nan
    return r


def chain(name, L=400, R=25, focal_plane=0, **kw):
    """The reference's diagnostic layouts as op lists (rtm_solver.py:197-286,376-422)."""
    if name == "shadow_single":                                          # rtm_solver.py:197-203
        return [("travel", 3 * L / 4 - focal_plane), ("circ_ap", R), ("lens", L / 2, L / 2), ("travel", 3 * L / 2)]
    if name in ("shadow_two", "interf_two"):                             # rtm_solver.py:205-214,376-422
        return [("travel", L - focal_plane), ("circ_ap", R), ("lens", L / 2, L / 2), ("travel", L * 2),
                ("circ_ap", R), ("lens", L / 2, L / 2), ("travel", L)]
    if name == "shadow_single_exp":                                      # rtm_solver.py:216-222
        return [("travel", L), ("circ_ap", R), ("lens", L / 2, L / 2), ("travel", kw.get("detL", 400))]
    if name in ("schlieren_DF", "schlieren_LF"):                         # rtm_solver.py:231-267
        mid = ("circ_stop" if name.endswith("DF") else "circ_ap", kw.get("R_stop", 1))
        return [("travel", L - focal_plane), ("circ_ap", R), ("lens", L, L), ("travel", L), mid,
                ("travel", L), ("circ_ap", R), ("lens", L, L), ("travel", L)]
    if name == "refracto_incoherent":                                    # rtm_solver.py:276-286
        return [("travel", 3 * L / 4 - focal_plane), ("circ_ap", R), ("lens", L / 2, L / 2), ("travel", 3 * L / 2),
                ("rect_ap", 15, 30), ("circ_ap", R), ("lens", L / 3, L / 2), ("travel", L)]
    if name == "refracto_coherent":                                      # rtm_solver.py:288-331
        # quirk kept: the field is NOT advanced across the middle travel (rtm_solver.py:308-314 take
        # r5 - r4 *after* the in-place aperture, i.e. zero) -> "travel_noE"
        return [("travel", 3 * L / 4 - focal_plane), ("circ_ap", R), ("lens", L / 2, L / 2), ("travel_noE", 3 * L / 2),
                ("circ_ap", R), ("lens", L / 3, L / 2), ("travel", L)]
    if name == "schlieren_knife":       # SURVEY.md 8d-C4 (knife edge as the dark-field stop; upstream's own knife-edge train,
        # rtm_solver-louis.py:375-391, is pinned as an explicit op list in tests: golden g13)
        return [("travel", L), ("circ_ap", R), ("lens", L, L), ("travel", L),
                ("knife", kw.get("offset", 0.0), kw.get("axis", 2), kw.get("direction", 1)),
                ("travel", L), ("circ_ap", R), ("lens", L, L), ("travel", L)]
    raise ValueError(name)


def run_chain(rf_m, ops, E=None, wl=None):
    """rf_m: (4,N) exit rays in metres.  Applies ``m_to_mm`` (rtm_solver.py:153) then the ops.  With
    ``E`` (2,N complex) the field picks up exp(i k sqrt(dx^2+dy^2)) across every *matrix* element, as in
    rtm_solver.py:380-421 (positions in mm, k = 2 pi / wl with wl in metres -- the reference's unit mix)."""
    r = m_to_mm(rf_m)
    if E is not None:
        E = np.array(E, dtype=complex, copy=True)
        k = 2 * np.pi / wl
    for op in ops:
        r_new = apply_op(r, op)
        if E is not None and op[0] in ("travel", "lens"):
            dx, dy = r_new[0] - r[0], r_new[2] - r[2]
            E = E * np.exp(1.0j * k * np.sqrt(dx ** 2 + dy ** 2))
        r = r_new
    return (r, E) if E is not None else r


def coherent_solve_current(rf_m, E, wl, L=400, R=25, focal_plane=0):
    """``Refractometry.coherent_solve`` of the CURRENT generation (src/simulator/diagnostics.py:505-524), which differs
    from the legacy one (rtm_solver.py:288-331 == chain("refracto_coherent")): the first aperture is applied to ``r0``,
    not to the rays carried to the first lens, so the positions never make the ``3L/4 - focal_plane`` leg; the field is
    advanced by the length of that leg all the same (``propagate_E(r2, r1)``) and, this time, across the middle travel.
    Pinned by tests/golden/g9_diagnostics.npz (the reference's own source run on NumPy arrays)."""
    k = 2 * np.pi / wl
    E = np.array(E, dtype=complex, copy=True)

    def advance(E, a, b):                                                # diagnostics.py:315-321
        return E * np.exp(1.0j * k * np.sqrt((a[0] - b[0]) ** 2 + (a[2] - b[2]) ** 2))

    r0 = m_to_mm(rf_m)
    r1 = apply_op(r0, ("travel", 3 * L / 4 - focal_plane))
    r2 = apply_op(r0, ("circ_ap", R))
    E[:, np.isnan(r2[0])] = np.nan
    E = advance(E, r2, r1)
    r3 = apply_op(r2, ("lens", L / 2, L / 2))
    E = advance(E, r3, r2)
    r4 = apply_op(r3, ("travel", 3 * L / 2))
    E = advance(E, r4, r3)
    r5 = apply_op(r4, ("circ_ap", R))
    E[:, np.isnan(r5[0])] = np.nan
    r6 = apply_op(r5, ("lens", L / 3, L / 2))
    E = advance(E, r6, r5)
    r7 = apply_op(r6, ("travel", L))
    return r7, advance(E, r7, r6)


def histogram(r, bin_scale=10, pix_x=3448, pix_y=2574, Lx=18, Ly=13.5):
    """rtm_solver.py:156-174: drop NaN rays, np.histogram2d on the detector rectangle, transpose."""
    x, y = r[0], r[2]
    x, y = x[~np.isnan(x)], y[~np.isnan(y)]
    H, _, _ = np.histogram2d(x, y, bins=[pix_x // bin_scale, pix_y // bin_scale],
                             range=[[-Lx / 2, Lx / 2], [-Ly / 2, Ly / 2]])
    return H.T


def interferogram_edges(bin_scale=1, pix_x=3448, pix_y=2574, Lx=18, Ly=13.5):
    """Bin edges of rtm_solver.py:436-437.  NB ``-Ly//2`` is ``(-Ly)//2``: Ly=13.5 gives [-7, 6]."""
    return (np.linspace(-Lx // 2, Lx // 2, pix_x // bin_scale),
            np.linspace(-Ly // 2, Ly // 2, pix_y // bin_scale))


def interferogram(r, E, bin_scale=1, pix_x=3448, pix_y=2574, Lx=18, Ly=13.5, return_planes=False):
    """rtm_solver.py:424-453 without the per-ray Python loop (np.add.at is the same unordered sum up to
    FP64 association; checked against the reference's loop in tests/test_oracle_golden.py)."""
    xb, yb = interferogram_edges(bin_scale, pix_x, pix_y, Lx, Ly)
    ax = np.zeros((len(yb) - 1, len(xb) - 1), dtype=complex)
    ay = np.zeros_like(ax)
    xi = np.digitize(r[0], xb) - 1
    yi = np.digitize(r[2], yb) - 1
    ok = (xi >= 0) & (xi < ax.shape[1]) & (yi >= 0) & (yi < ax.shape[0])
    np.add.at(ax, (yi[ok], xi[ok]), E[0, ok])
    np.add.at(ay, (yi[ok], xi[ok]), E[1, ok])
    H = np.sqrt(np.real(ax) ** 2 + np.real(ay) ** 2)
    return (H, ax, ay) if return_planes else H


def interfere_ref_beam(rf_m, E, n_fringes=10, deg=20):
    """Reference-beam term of the *JAX-generation* API (/root/reference/src/simulator/diagnostics.py:
    559-581).  Pinned by tests/golden/g9_diagnostics.npz: that file's own source executed with a NumPy stand-in for
    jax.numpy (oracle/gen_golden.py::import_diagnostics).
    Note the reference evaluates it on ``self.rf`` i.e. exit positions in METRES (diagnostics.py:579)."""
    if deg >= 45:
        deg = -abs(deg - 90)
    rad = deg * np.pi / 180
    yw = np.arctan(rad)
    xw = np.sqrt(1 - yw ** 2)
    E = np.array(E, dtype=complex, copy=True)
    E[1] = E[1] + np.exp(2 * n_fringes / 3 * 1.0j * (xw * rf_m[0] + yw * rf_m[2]))
    return E


# --------------------------------------------------------------------------------------------------
# Wave-optics step                                     (/root/reference/src/simulator/fresnel_integral.py)
# --------------------------------------------------------------------------------------------------
# Pinned: tests/golden/g7_fresnel.npz is produced by the reference module itself (NumPy/SciPy only).

def scatter_to_grid(px, py, values, x, y, fill_value=0.0):
    """fresnel_integral.py:71-77: piecewise-linear interpolation of scattered samples on their Delaunay
    triangulation (``scipy.interpolate.LinearNDInterpolator``, the reference's own call), evaluated on
    ``np.meshgrid(x, y)`` -> (len(y), len(x)); ``fill_value`` outside the convex hull."""
    from scipy.interpolate import LinearNDInterpolator
    XX, YY = np.meshgrid(x, y)
    return LinearNDInterpolator((px, py), values, fill_value=fill_value)((XX, YY))


def tukey_window(M, alpha=0.4):
    """``scipy.signal.windows.tukey(M, alpha)`` (symmetric) written out: cosine tapers of width
    floor(alpha (M-1) / 2) + 1 samples at both ends, 1 between."""
    n = np.arange(M, dtype=float)
    if alpha <= 0:
        return np.ones(M)
    if alpha >= 1:
        return 0.5 - 0.5 * np.cos(2 * np.pi * n / (M - 1))
    width = int(np.floor(alpha * (M - 1) / 2.0))
    w = np.ones(M)
    head, tail = n[:width + 1], n[M - width - 1:]
    w[:width + 1] = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * head / alpha / (M - 1))))
    w[M - width - 1:] = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * tail / alpha / (M - 1))))
    return w


def reflect_index(i, n):
    """Source index of ``np.pad(..., mode='reflect')`` for (possibly far) out-of-range i: period 2(n-1), no
    repeated edge sample."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    m = np.mod(i, p)
    return np.where(m < n, m, p - m)


def fresnel_prepare(U0, pad_factor=2, alpha=0.4):
    """fresnel_integral.py:7-24: reflect-pad by pad_factor x size on every side, then a separable Tukey window."""
    n0, n1 = U0.shape
    i0 = reflect_index(np.arange(-n0 * pad_factor, n0 * (1 + pad_factor)), n0)
    i1 = reflect_index(np.arange(-n1 * pad_factor, n1 * (1 + pad_factor)), n1)
    return U0[np.ix_(i0, i1)] * np.outer(tukey_window(len(i0), alpha), tukey_window(len(i1), alpha))


def fresnel_propagate(U0_prepared, L, wavelength, z, original_shape, pad_factor=2, lanex_fwhm_m=None):
    """fresnel_integral.py:27-59: transfer function exp(-i pi lambda z (fx^2 + fy^2)) on the padded grid (sample
    spacing L / original size), optional Gaussian PSF, factor exp(i k z) / (i lambda z), crop to the original window."""
    n0, n1 = original_shape
    f0 = np.fft.fftfreq(U0_prepared.shape[0], d=L[0] / n0)
    f1 = np.fft.fftfreq(U0_prepared.shape[1], d=L[1] / n1)
    F2 = f0[:, None] ** 2 + f1[None, :] ** 2
    spec = np.fft.fft2(U0_prepared) * np.exp(-1j * np.pi * wavelength * z * F2)
    if lanex_fwhm_m is not None and lanex_fwhm_m > 0:
        sigma = lanex_fwhm_m / (2 * np.sqrt(2 * np.log(2)))
        spec = spec * np.exp(-2 * (np.pi * sigma) ** 2 * F2)
    out = np.fft.ifft2(spec) * np.exp(1j * (2 * np.pi / wavelength) * z) / (1j * wavelength * z)
    return out[n0 * pad_factor:n0 * (pad_factor + 1), n1 * pad_factor:n1 * (pad_factor + 1)]


def fresnel(lwl, x, y, x_length, y_length, rays, amplitudes, phases, z, pad_factor=2):
    """fresnel_integral.py:61-93 ``propagate``: rays (4,N) rows 0 / 2 are the sample positions."""
    ph = scatter_to_grid(rays[0], rays[2], phases, x, y)
    am = scatter_to_grid(rays[0], rays[2], amplitudes, x, y)
    U0 = am * np.exp(-1j * ph)
    return fresnel_propagate(fresnel_prepare(U0, pad_factor), (x_length, y_length), lwl, z, U0.shape, pad_factor)
