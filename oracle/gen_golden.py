"""Generate tests/golden/*.npz by running the REAL reference in the build container.

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/
    python oracle/gen_golden.py diagnostics   # only g9-g12 (current generation, below)

g1-g8: the legacy generation (NumPy/SciPy), imported as it is.  g9-g12: the current generation's files
(src/simulator/{diagnostics,beam,utils,domain,propagator}.py), whose sources are executed UNMODIFIED with a NumPy stand-in
registered as ``jax.numpy`` (``jnp_shim`` / ``simulator_stubs`` below): those files use jax.numpy as an array library on the
paths exercised (no jit / vmap / lax), so what is pinned is their arithmetic in float64 (= ``jax_enable_x64``); XLA's
elementary functions may differ from NumPy's in the last ulp.  ``propagator.solve`` (diffrax) cannot be run this way.

The reference (/root/reference/src/solvers-legacy/{full_solver,rtm_solver}.py and
/root/reference/src/field_generator/gaussian3D.py) is imported unmodified with the two zero-source-change
shims of SURVEY.md section 8c:
  1. ``full_solver.omega_pe = full_solver.ScalarDomain.omega_pe``  (module global missing -> NameError at
     full_solver.py:273 whenever phaseshift=True)
  2. empty ``matplotlib`` / ``matplotlib.pyplot`` modules registered before ``import rtm_solver``
     (rtm_solver.py:8-9 import them at top level; unused by the code paths exercised).
The fixtures travel to the GPU box; /root/reference does not.  Inputs are stored next to outputs so the
tests never need to regenerate anything with a RNG.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, os.path.join(REF, "solvers-legacy"))
    sys.path.insert(0, os.path.join(REF, "field_generator"))
    import full_solver as fs
    import rtm_solver as rtm
    import gaussian3D as g3
    fs.omega_pe = fs.ScalarDomain.omega_pe
    return fs, rtm, g3


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def joint_log(fs, dom, s0):
    """(h, error_norm) of every attempted step of the shipped joint solve (instrumented scipy RK45)."""
    from scipy.integrate import RK45
    s = RK45(lambda t, y: fs.dsdt(t, y, dom), 0.0, s0.ravel().copy(), np.sqrt(8.0) * dom.extent / fs.c)
    orig, log = s._estimate_error_norm, []
    def est(K, h, scale):
        e = orig(K, h, scale)
        log.append((h, e))
        return e
    s._estimate_error_norm = est
    while s.status == "running":
        s.step()
    return np.array(log).T


def axes(lengths, dims):
    return [np.linspace(-L / 2, L / 2, n) for L, n in zip(lengths, dims)]


def gaussian_column(x, y, z, ne0=1e24, LR=1e-3):
    """Formula of minimal_solver.test_lens (minimal_solver.py:192-201), loaded via external_ne."""
    XX, YY, _ = np.meshgrid(x, y, z, indexing="ij")
    return ne0 * np.exp(-(XX ** 2 + YY ** 2) / LR ** 2)


def probe_states(rng, n, lengths, frac_out=0.15):
    """Random 9-vectors incl. out-of-box, on-face and exact-node positions."""
    from scipy.constants import c
    half = np.array(lengths)[:, None] / 2
    pos = (rng.random((3, n)) * 2 - 1) * half
    k = int(frac_out * n)
    pos[:, :k] *= 1.0 + 0.2 * rng.random((3, k))            # some outside
    v = rng.standard_normal((3, n)) * 0.05 * c
    v[2] += c
    s = np.zeros((9, n))
    s[:3], s[3:6] = pos, v
    s[6] = 1.0 + 0.1 * rng.random(n)
    s[7] = rng.random(n)
    return s


def fresnel_fixture():
    """G7: the reference's wave-optics step, src/simulator/fresnel_integral.py (NumPy/SciPy only, imported unmodified):
    scattered rays -> LinearNDInterpolator grids -> reflect pad + Tukey window -> Fresnel transfer function."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fresnel_integral", os.path.join(REF, "simulator", "fresnel_integral.py"))
    fi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fi)
    from scipy.interpolate import LinearNDInterpolator as LND
    rng = np.random.default_rng(21)
    N, nx, ny = 6000, 96, 72
    Lx, Ly = 18e-3, 13.5e-3
    r0 = np.zeros((4, N))
    r0[0] = rng.uniform(-0.48 * Lx, 0.48 * Lx, N)            # the hull stops short of the grid: fill_value region
    r0[2] = rng.uniform(-0.47 * Ly, 0.47 * Ly, N)
    r0[1], r0[3] = rng.normal(0, 1e-3, (2, N))
    amp = 1.0 + 0.3 * np.cos(2 * np.pi * r0[0] / 4e-3) * np.exp(-(r0[2] / 5e-3) ** 2)
    phase = 40.0 * np.exp(-(r0[0] ** 2 + r0[2] ** 2) / (3e-3) ** 2) + 0.05 * rng.normal(size=N)
    x, y = np.linspace(-Lx / 2, Lx / 2, nx), np.linspace(-Ly / 2, Ly / 2, ny)
    lwl, z = 1064e-9, 0.3
    XX, YY = np.meshgrid(x, y)
    g = dict(r0=r0, amp=amp, phase=phase, x=x, y=y, Lx=Lx, Ly=Ly, lwl=lwl, z=z)
    g["phase_grid"] = LND((r0[0], r0[2]), phase, fill_value=0.0)((XX, YY))          # fresnel_integral.py:71-77
    g["amp_grid"] = LND((r0[0], r0[2]), amp, fill_value=0.0)((XX, YY))
    U0 = g["amp_grid"] * np.exp(-1j * g["phase_grid"])
    for pf in (2, 1):
        prep = fi.prepare_field_for_propagation(U0, pad_factor=pf)
        g["prep_pf%d_sub" % pf] = prep[::7, ::5]                                     # subsample: keeps the fixture small
        g["out_pf%d" % pf] = fi.propagate(lwl, x, y, Lx, Ly, r0, amp, phase, z, pad_factor=pf)
    g["out_lanex"] = fi.fresnel_propagate(fi.prepare_field_for_propagation(U0), (Lx, Ly), lwl, z, U0.shape,
                                          lanex_fwhm_m=150e-6)
    np.savez_compressed(os.path.join(OUT, "g7_fresnel.npz"), **g)


def minimal_fixture():
    """G8: the 6-component generation of the solver (src/solvers-legacy/minimal_solver.py): float64 axes and gradients,
    ne_max clamp, RMS error norm over 6N components, its own integration span.  Axes are dyadic (exact in float32) so
    that the only representation difference to the float32 field layout is the rounding of the gradient values."""
    sys.path.insert(0, os.path.join(REF, "solvers-legacy"))
    import minimal_solver as ms
    from scipy.integrate import solve_ivp
    n = 33
    x = (np.arange(n) - n // 2) * 2.0 ** -13           # +-1.95 mm, spacing 0.122 mm
    y = (np.arange(n) - n // 2) * 2.0 ** -13
    z = (np.arange(n) - n // 2) * 2.0 ** -12           # +-3.9 mm
    g = dict(x=x, y=y, z=z, lwl=1064e-9, ne_max=0.02)
    for tag, make in (("lens", lambda d: d.test_lens(n_e0=3e25, LR=8e-4)),):
        dom = ms.ScalarDomain(x, y, z, "z")
        make(dom)
        g[tag + "_ne"] = dom.ne.copy()
        dom.calc_dndr(lwl=g["lwl"], ne_max=g["ne_max"])      # clamps ne_nc in place (and therefore dom.ne / nc)
        np.random.seed(12)
        quiet(dom.init_beam, 96, 1.5e-3, 1e-4)
        g[tag + "_s0"] = dom.s0.copy()
        rf = quiet(dom.solve)
        g[tag + "_sf"], g[tag + "_rf"] = dom.sf.copy(), rf
        g[tag + "_t_end"] = np.sqrt(dom.extent_x ** 2 + dom.extent_y ** 2 * dom.extent_z ** 2) / ms.c     # minimal_solver.py:321
        sol = solve_ivp(lambda t, yv: ms.dsdt(t, yv, dom), [0, g[tag + "_t_end"]], dom.s0.flatten(), t_eval=[0, g[tag + "_t_end"]])
        assert np.array_equal(sol.y[:, -1].reshape(6, -1), dom.sf)
        g[tag + "_nfev"] = sol.nfev
        g[tag + "_dndx"], g[tag + "_dndy"], g[tag + "_dndz"] = dom.dndx, dom.dndy, dom.dndz
        probe = np.concatenate([dom.s0, dom.sf], axis=1)
        g[tag + "_probe"] = probe
        g[tag + "_dsdt"] = ms.dsdt(0.0, probe.flatten(), dom).reshape(6, -1)
    np.savez_compressed(os.path.join(OUT, "g8_minimal.npz"), **g)


def jnp_shim():
    """A ``jax.numpy`` stand-in made of NumPy, for running src/simulator/diagnostics.py UNMODIFIED where jax cannot be
    installed.  That file uses jax.numpy purely as an array library (no jit / vmap / lax / random in any code path
    exercised): 18 NumPy-named functions plus the functional update ``a.at[idx].set(v)``.  The stand-in maps the
    functions to NumPy's float64 ones (== jax with ``jax_enable_x64``, config.py:129-131; elementary functions may differ
    from XLA's in the last ulp) and gives arrays an ``at`` that returns an updated COPY.  ``array == None`` is False, as it
    is for a jax array (diagnostics.py:569 relies on it)."""
    class _Idx:
        def __init__(self, a, idx):
            self.a, self.idx = a, idx

        def set(self, v):
            out = np.array(self.a, copy=True).view(JArray)
            out[self.idx] = v
            return out

    class _At:
        def __init__(self, a):
            self.a = a

        def __getitem__(self, idx):
            return _Idx(self.a, idx)

    class JArray(np.ndarray):
        @property
        def at(self):
            return _At(self)

        def __eq__(self, other):
            return False if other is None else np.ndarray.__eq__(self, other)

        def __iadd__(self, other):                          # jax arrays are immutable: ``a += b`` rebinds, and may broadcast a up
            return self + other

        def __imul__(self, other):
            return self * other

        __hash__ = None

    def wrap(v):
        if isinstance(v, tuple):
            return tuple(wrap(u) for u in v)
        return v.view(JArray) if isinstance(v, np.ndarray) else v

    mod = types.ModuleType("jax.numpy")
    for name in ("abs", "arctan", "array", "asarray", "copy", "diag", "digitize", "exp", "histogram2d", "isnan", "linspace",
                 "matmul", "real", "rot90", "sqrt", "tanh", "zeros", "cos", "sin", "shape", "max", "meshgrid", "pad", "stack", "zeros_like",
                 "sum", "round", "maximum", "expand_dims", "concatenate", "append", "floor", "log10", "gradient", "log", "ones", "power",
                 "ravel", "reshape", "where", "broadcast_arrays", "empty", "searchsorted"):
        setattr(mod, name, (lambda f: lambda *a, **k: wrap(f(*a, **k)))(getattr(np, name)))
    mod.nan, mod.pi, mod.int32, mod.float32 = np.nan, np.pi, np.int32, np.float32
    return mod


@contextlib.contextmanager
def simulator_stubs():
    """What src/simulator/*.py needs to be importable where jax is absent, registered in sys.modules for the duration:
    ``jax.numpy`` -> ``jnp_shim()``, empty matplotlib modules, ``equinox.Module`` as a plain base class, ``jax.Array``,
    ``jax.lib.xla_bridge.get_backend().platform == 'cpu'`` (domain.py:141-142), name-only placeholders for the jax internals
    utils.py:115-118 imports for its own interpolator (not exercised), and an empty ``propagator`` (only
    ``Interferometry.bkg``, not exercised, uses it).  The sibling modules (utils, printing, fresnel_integral) are the real ones."""
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    jax = types.ModuleType("jax")
    jax.__path__ = []
    jax.numpy = jnp_shim()
    jax.Array = np.ndarray
    prop = types.ModuleType("propagator")
    prop.ray_to_Jonesvector = None
    eqx = types.ModuleType("equinox")
    eqx.Module = type("Module", (), {})
    bridge = types.ModuleType("jax.lib.xla_bridge")
    bridge.get_backend = lambda: types.SimpleNamespace(platform="cpu")
    lib = types.ModuleType("jax.lib")
    lib.__path__, lib.xla_bridge = [], bridge
    stubs = {"jax": jax, "jax.numpy": jax.numpy, "propagator": prop, "equinox": eqx, "jax.lib": lib, "jax.lib.xla_bridge": bridge}
    # utils.py:115-118 takes these from jax internals for its own RegularGridInterpolator: the same NumPy stand-ins
    internals = {"jax._src": {}, "jax._src.dtypes": {"can_cast": np.can_cast}, "jax._src.tree_util": {"register_pytree_node": None},
                 "jax._src.numpy": {k: getattr(jax.numpy, k) for k in ("asarray", "broadcast_arrays", "empty", "searchsorted", "where", "zeros")},
                 "jax._src.numpy.util": {"check_arraylike": lambda *a, **k: None,
                                         "promote_dtypes_inexact": lambda *a: [jax.numpy.asarray(v, dtype=np.result_type(v, np.float32)) for v in a]}}
    for name, attrs in internals.items():
        m = types.ModuleType(name)
        m.__path__ = []
        for a, v in attrs.items():
            setattr(m, a, v)
        stubs[name] = m
    stubs["jax._src"].dtypes = stubs["jax._src.dtypes"]
    names = tuple(stubs) + ("utils", "fresnel_integral", "printing")
    saved = {k: sys.modules.pop(k, None) for k in names}
    sys.modules.update(stubs)
    sys.path.insert(0, os.path.join(REF, "simulator"))
    try:
        yield
    finally:
        sys.path.pop(0)
        for k in names:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]


def import_simulator(module):
    """``src/simulator/<module>.py`` executed from its own source under ``simulator_stubs()``."""
    import importlib.util
    with simulator_stubs():
        spec = importlib.util.spec_from_file_location("ref_" + module, os.path.join(REF, "simulator", module + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return mod


def import_diagnostics():
    return import_simulator("diagnostics")


def domain_fixture():
    """G11: ``ScalarDomain`` of the current generation (src/simulator/domain.py:11-451) executed from its own source: the
    float32 axes of a non-cubic grid and the four named profiles, evaluated upstream on the float32-rounded mesh."""
    dm = import_simulator("domain")
    lengths, dims = (8e-3, 6e-3, 10e-3), (14, 11, 9)
    g = dict(lengths=np.array(lengths), dims=np.array(dims))
    with simulator_stubs():                                        # the constructor imports jax.lib when it is called
        for name in ("test_null", "test_slab", "test_linear_cos", "test_exponential_cos"):
            d = quiet(dm.ScalarDomain, lengths, dims, ne_type=name)
            g[name] = np.asarray(d.ne)
        g["x"], g["y"], g["z"] = (np.asarray(v) for v in (d.x, d.y, d.z))
        d = quiet(dm.ScalarDomain, 5e-3, 7)                        # scalar arguments; the mesh is kept for the caller
        g["cube_x"], g["cube_XX"] = np.asarray(d.x), np.asarray(d.XX)
    np.savez_compressed(os.path.join(OUT, "g11_domain.npz"), **g)


def propagator_fixture():
    """G12: the array-level functions of the current generation's propagator.py executed from their own source -- the RHS
    ``dsdt`` (:94-175; gradient of ne / (3.142e-4 omega^2) taken at every call, utils.RegularGridInterpolator :124-214),
    ``ray_to_Jonesvector`` (:178-298) and ``back_propogate`` (:300-349) -- on the grid and probe states of G1 with the
    float32 axes a current-generation ScalarDomain holds.  (``solve`` itself is diffrax / jit code and cannot run here.)"""
    pr = import_simulator("propagator")
    g1 = np.load(os.path.join(OUT, "g1_rhs.npz"))
    x, y, z = (np.float32(g1[k]) for k in "xyz")
    lwl = float(g1["lwl"])
    omega = 2 * np.pi * pr.c / lwl                                           # propagator.py: omega = 2 pi c / lwl
    s = g1["s"]
    g = dict(omega=omega)
    with simulator_stubs():
        g["dsdt"] = np.asarray(pr.dsdt(0.0, s.ravel().copy(), False, False, False, False, g1["ne"], None, None, None, x, y, z,
                                       omega, None, None, None)).reshape(9, -1)
        rng = np.random.default_rng(5)
        sf = np.zeros((9, 257))
        sf[:3] = rng.uniform(-4e-3, 4e-3, (3, 257))
        sf[3:6] = rng.normal(0, 2e6, (3, 257))
        sf[6], sf[7], sf[8] = rng.uniform(0.2, 1, 257), rng.uniform(0, 300, 257), rng.uniform(-0.3, 0.3, 257)
        g["sf"] = sf
        for pd, col in (("x", 3), ("y", 4), ("z", 5)):
            st = sf.copy()
            st[col] = pr.c * (1 - rng.uniform(0, 1e-4, 257))
            g["sf_" + pd] = st
            for keep in (False, True):
                rp, rj = pr.ray_to_Jonesvector(st.copy(), 5e-3, probing_direction=pd, keep_current_plane=keep, return_E=True)
                g["rtj_%s_%d_p" % (pd, keep)], g["rtj_%s_%d_J" % (pd, keep)] = np.asarray(rp), np.asarray(rj)
            g["bp_" + pd] = np.asarray(pr.back_propogate(st.copy().view(type(pr.jnp.zeros(1))), 5e-3, pd))
        # NRL inverse-bremsstrahlung rate and refractive index as the current generation codes them (:23-64)
        ne3 = 10.0 ** rng.uniform(22, 27.2, (6, 5, 4))                       # up to above the critical density (omega_pe > omega branch)
        Te3, Z3 = 10.0 ** rng.uniform(0, 3.5, (6, 5, 4)), rng.uniform(1, 30, (6, 5, 4))
        as_j = lambda a: a.view(type(pr.jnp.zeros(1)))
        g["k_ne"], g["k_Te"], g["k_Z"] = ne3, Te3, Z3
        g["kappa"] = np.asarray(pr.kappa(as_j(ne3.copy()), as_j(Te3.copy()), as_j(Z3.copy()), omega))
        with np.errstate(invalid="ignore"):
            g["n_refrac"] = np.asarray(pr.n_refrac(as_j(ne3.copy()), omega))
    np.savez_compressed(os.path.join(OUT, "g12_propagator.npz"), **g)


def current_solve_fixture():
    """G14: the current generation end to end on its SciPy path -- ``domain.ScalarDomain(ne_type=...)`` ->
    ``beam.Beam`` -> ``propagator.solve(..., parallelise=False)`` (``solve_ivp`` RK45 over all rays jointly with the
    current ``dsdt``, propagator.py:466-474) -> ``ray_to_Jonesvector`` -- every file executed from its own source."""
    dm, bm, pr = import_simulator("domain"), import_simulator("beam"), import_simulator("propagator")
    lengths, dims = (6e-3, 6e-3, 8e-3), (28, 24, 36)
    g = dict(lengths=np.array(lengths), dims=np.array(dims), lwl=1064e-9)
    with simulator_stubs():
        for tag, pd, ext in (("z", "z", 4e-3), ("x", "x", 3e-3)):
            dom = quiet(dm.ScalarDomain, lengths, dims, ne_type="test_exponential_cos", probing_direction=pd)
            np.random.seed(21)
            beam = quiet(bm.Beam, 160, 1.5e-3, 1e-4, ext, probing_direction=pd)
            s0 = np.asarray(beam.s0).copy()
            rf, Jf, _ = quiet(pr.solve, beam.s0, dom, ext, return_E=True, parallelise=False, lwl=1064e-9)
            g[tag + "_s0"], g[tag + "_rf"], g[tag + "_Jf"], g[tag + "_extent"] = s0, np.asarray(rf), np.asarray(Jf), ext
        g["ne"] = np.asarray(dom.ne)
    np.savez_compressed(os.path.join(OUT, "g14_current_solve.npz"), **g)


def louis_fixture():
    """G13: src/solvers-legacy/rtm_solver-louis.py (sympy-lambdified composite matrices; the one place upstream uses a knife
    edge inside a diagnostic, SchlierenRays.solve :375-391), imported unmodified by path (its file name is not a module
    name).  Rays are given in mm, as that file expects."""
    import importlib.util
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("rtm_solver_louis", os.path.join(REF, "solvers-legacy", "rtm_solver-louis.py"))
    lo = importlib.util.module_from_spec(spec)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                     # "is" with a literal (:133)
        spec.loader.exec_module(lo)
    rng = np.random.default_rng(13)
    N = 4000
    r0 = np.zeros((4, N))
    r0[0], r0[2] = rng.uniform(-6, 6, (2, N))                                # mm
    r0[1], r0[3] = rng.normal(0, 3e-3, (2, N))
    r0[1, :300], r0[3, :300] = rng.normal(0, 6e-2, (2, 300))                 # rays the lens apertures reject
    E = (rng.normal(size=(2, N)) + 1j * rng.normal(size=(2, N))) / np.sqrt(2)
    g = dict(r0=r0, E=E, L=400.0, R=25.0, displacement=7.5, wl=532e-9)
    for cls, tag, kw in ((lo.RefractometerRays, "refractometer", {}), (lo.ShadowgraphyRays, "shadowgraphy", {"displacement": 7.5}),
                         (lo.SchlierenRays, "schlieren", {})):
        d = cls(r0.copy(), L=400, R=25)
        d.solve(**kw)
        g[tag + "_rf"] = d.rf
        d.histogram(bin_scale=24)
        g[tag + "_H"] = d.H
    d = lo.InterferometerRays(r0.copy(), E=E.copy(), L=400, R=25)
    d.solve(wl=532e-9)
    g["interferometer_rf"], g["interferometer_rE"] = d.rf, d.rE
    np.savez_compressed(os.path.join(OUT, "g13_louis.npz"), **g)


def beam_fixture():
    """G10: ``Beam`` of the current generation (src/simulator/beam.py:7-303 + utils.py:8-24, NumPy's global RNG) executed from
    its own source: every beam type that upstream can construct x probing direction, unseeded (np.random.seed set by the
    caller) and ``seeded=True`` (utils re-seeds with 0 before every draw, so t, u and chi share one stream start)."""
    bm = import_simulator("beam")
    g = {}
    # ('linear' and 'even' cannot be constructed upstream: beam.py:291 deletes a name 'linear' never binds, :222 ranges over a float)
    for bt, size in (("circular", 4e-3), ("square", 3e-3), ("rectangular", (1e-3, 2.5e-3)), ("rect_trackers", (2e-3, 0.5e-3))):
        for pd in ("x", "y", "z"):
            for seeded in (False, True):
                np.random.seed(17)
                b = quiet(bm.Beam, 96, size, 2e-4, 6e-3, probing_direction=pd, beam_type=bt, seeded=seeded)
                g["%s_%s_%d" % (bt, pd, seeded)] = np.asarray(b.s0)
    np.savez_compressed(os.path.join(OUT, "g10_beam.npz"), **g)


def diagnostics_fixture():
    """G9: the CURRENT generation's optics and detector code (src/simulator/diagnostics.py:122-640) executed from its own
    source through ``import_diagnostics``: every ``*_solve`` layout, ``propagate_E``, the reference beam
    (``interfere_ref_beam`` on exit rays in metres), ``histogram`` and the per-ray ``histogram_legacy`` loop."""
    dg = import_diagnostics()
    rng = np.random.default_rng(9)
    N, lwl = 3000, 1064e-9
    rf = np.zeros((4, N))
    rf[0], rf[2] = rng.uniform(-8e-3, 8e-3, (2, N))                   # exit plane, metres
    rf[1], rf[3] = rng.normal(0, 4e-3, (2, N))
    wide = rng.choice(N, 400, replace=False)                          # rays the lens apertures / stops reject
    rf[1, wide], rf[3, wide] = rng.normal(0, 5e-2, (2, 400))
    rf[1, :40], rf[3, :40] = rng.normal(0, 2e-6, (2, 40))             # rays the dark-field stop rejects
    rf[:, 40:44] = np.nan                                             # rays lost before the exit plane
    Jf = (rng.normal(size=(2, N)) + 1j * rng.normal(size=(2, N))) / np.sqrt(2)
    g = dict(rf=rf, Jf=Jf, lwl=lwl, L=400.0, R=25.0, focal_plane=3.0, bin_scale=24)
    kw = dict(focal_plane=3.0, L=400, R=25)

    def image(d, which, **k):
        quiet(getattr(d, which), bin_scale=24, **k)
        return np.asarray(d.H)

    for cls, meth, args in (("Shadowgraphy", "single_lens_solve", {}), ("Shadowgraphy", "two_lens_solve", {}),
                            ("Schlieren", "DF_solve", {"R": 1}), ("Schlieren", "LF_solve", {"R": 1}),
                            ("Refractometry", "incoherent_solve", {})):
        d = getattr(dg, cls)(lwl, rf.copy(), **kw)
        getattr(d, meth)(**args)
        g[meth + "_rf"], g[meth + "_H"] = np.asarray(d.rf), image(d, "histogram")
    d = dg.Refractometry(lwl, rf.copy(), Jf.copy(), **kw)
    d.coherent_solve()
    g["coherent_solve_rf"], g["coherent_solve_Jf"], g["coherent_solve_H"] = np.asarray(d.rf), np.asarray(d.Jf), image(d, "refractogram")
    d = dg.Refractometry(lwl, rf.copy(), Jf.copy(), focal_plane=0, L=300, R=6)      # the aperture on r0 rejects rays
    d.coherent_solve()
    g["coherent_R6_rf"], g["coherent_R6_Jf"], g["coherent_R6_H"] = np.asarray(d.rf), np.asarray(d.Jf), image(d, "refractogram")
    d = dg.Interferometry(lwl, rf.copy(), Jf.copy(), **kw)
    d.interfere_ref_beam(7, 60)                                       # deg >= 45 branch
    g["ref_beam_7_60_Jf"] = np.asarray(d.Jf)
    d = dg.Interferometry(lwl, rf.copy(), Jf.copy(), **kw)
    d.two_lens_solve()
    g["interf_rf"], g["interf_Jf"], g["interf_H"] = np.asarray(d.rf), np.asarray(d.Jf), image(d, "interferogram")
    np.savez_compressed(os.path.join(OUT, "g9_diagnostics.npz"), **g)


def reference_fixture():
    """The one binary fixture the reference itself holds: evaluation/sergio_testing/integratedPy.npy = ne.sum(axis=2) of
    test_linear_cos(s1=-1, s2=1, n_e0=1e26, Ly=5e-3) on the 100 x 1000 x 100 grid of sergio_testing/notebook.ipynb cells
    7-8.  Copied byte for byte next to the grid parameters that produced it."""
    import shutil
    shutil.copyfile("/root/reference/evaluation/sergio_testing/integratedPy.npy", os.path.join(OUT, "integratedPy.npy"))


def main():
    os.makedirs(OUT, exist_ok=True)
    fs, rtm, g3 = import_reference()
    lwl = 1064e-9

    # ---------------------------------------------------------------- G1: RHS (L0) on a non-cubic grid
    lengths, dims, extent = (10e-3, 8e-3, 20e-3), (24, 20, 28), 10e-3
    x, y, z = axes(lengths, dims)
    ne = gaussian_column(x, y, z) * (1 + 0.3 * np.cos(2 * np.pi * z / 7e-3))[None, None, :]
    rng = np.random.default_rng(11)
    s = probe_states(rng, 4096, lengths)
    # exact nodes / faces / the rounded-float32 end points
    xf, yf, zf = np.float32(x), np.float32(y), np.float32(z)
    s[0, -8:] = np.float64(xf[[0, -1, 3, 3, 0, -1, 5, 7]])
    s[1, -8:] = np.float64(yf[[0, -1, 4, 0, -1, 2, 0, 19]])
    s[2, -8:] = np.float64(zf[[0, -1, 5, 27, 0, 13, 27, 27]])
    s[2, -16:-8] = -extent                                   # the beam's start plane (outside f32 z[0]?)
    out = {}
    for ph in (False, True):
        dom = fs.ScalarDomain(x, y, z, extent, phaseshift=ph)
        dom.external_ne(ne)
        dom.calc_dndr(lwl)
        out["dsdt_phase%d" % ph] = fs.dsdt(0.0, s.ravel().copy(), dom).reshape(9, -1)
    np.savez_compressed(os.path.join(OUT, "g1_rhs.npz"), x=x, y=y, z=z, ne=ne, extent=extent, lwl=lwl, s=s,
                        gradx=dom.dndx, grady=dom.dndy, gradz=dom.dndz, **out)

    # ---------------------------------------------------------------- G2: shipped joint RK45, exp-cos
    lengths, dims, extent = (10e-3, 10e-3, 20e-3), (40, 36, 48), 10e-3
    x, y, z = axes(lengths, dims)
    dom = fs.ScalarDomain(x, y, z, extent, phaseshift=True)
    dom.test_exponential_cos(n_e0=2e23, Ly=1e-3, s=-4e-3)     # evaluation/test_CoherentRefractogram.ipynb cell 2
    dom.calc_dndr(lwl)
    np.random.seed(0)
    s0 = fs.init_beam(384, 4e-3, 5e-5, extent, "circular", "z")
    rf, Jf = quiet(dom.solve, s0.copy(), return_E=True)
    g2 = dict(x=x, y=y, z=z, extent=extent, lwl=lwl, ne=dom.ne, s0=s0, sf=dom.sf, rf=rf, Jf=Jf,
              joint_log=joint_log(fs, dom, s0))

    # per-ray adaptive (== ScalarDomain.solve with Np=1 per ray), default and tight tolerances
    from scipy.integrate import solve_ivp
    sub = s0[:, :32]
    for tag, (rtol, atol) in {"def": (1e-3, 1e-6), "tight": (1e-7, 1e-9), "conv": (1e-11, 1e-13)}.items():
        if tag == "conv":
            sub = s0[:, :8]            # near-converged solution of the same RHS (slow): 8 rays
        sf1 = np.empty((9, sub.shape[1]))
        nfev = np.empty(sub.shape[1], dtype=np.int64)
        tt = np.linspace(0.0, np.sqrt(8.0) * extent / fs.c, 2)
        for i in range(sub.shape[1]):
            sol = solve_ivp(lambda t, yy: fs.dsdt(t, yy, dom), [0, tt[-1]], sub[:, i].copy(), t_eval=tt,
                            rtol=rtol, atol=atol)
            sf1[:, i], nfev[i] = sol.y[:, -1], sol.nfev
        g2["perray_sf_" + tag], g2["perray_nfev_" + tag] = sf1, nfev

    # fixed-step RK4 whose RHS is the reference dsdt
    def rk4(dom, s0, n_steps):
        h = np.sqrt(8.0) * dom.extent / fs.c / n_steps
        yv = s0.ravel().copy()
        f = lambda v: fs.dsdt(0.0, v, dom)
        for _ in range(n_steps):
            k1 = f(yv); k2 = f(yv + (0.5 * h) * k1); k3 = f(yv + (0.5 * h) * k2); k4 = f(yv + h * k3)
            yv = yv + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        return yv.reshape(9, -1)
    g2["rk4_nsteps"] = 160
    g2["rk4_sf"] = rk4(dom, s0[:, :128], 160)
    g2["rk4_rf"], g2["rk4_Jf"] = fs.ray_to_Jonesvector(g2["rk4_sf"], extent, probing_direction="z")
    np.savez_compressed(os.path.join(OUT, "g2_expcos.npz"), **g2)

    # ---------------------------------------------------------------- G3: turbulent field_generator grid
    np.random.seed(1)
    gen = g3.gaussian3D(lambda k: k ** (-11 / 3))
    f = gen.domain_fft(l_max=1, l_min=0.01, extent=5, res=16, factor=1)      # 32^3, examples turb_gen.py:36-50
    ne = 1e25 + 9e24 * f
    lengths, extent = (10e-3, 10e-3, 20e-3), 10e-3
    x, y, z = axes(lengths, ne.shape)
    dom = fs.ScalarDomain(x, y, z, extent)
    dom.external_ne(ne)
    dom.calc_dndr(lwl)
    np.random.seed(2)
    s0 = fs.init_beam(256, 4.5e-3, 5e-5, extent, "circular", "z")
    rf = quiet(dom.solve, s0.copy())
    g3d = dict(x=x, y=y, z=z, extent=extent, lwl=lwl, ne=ne, s0=s0, sf=dom.sf, rf=rf, joint_log=joint_log(fs, dom, s0),
               rk4_nsteps=200, rk4_sf=rk4(dom, s0[:, :128], 200))
    # other probing directions (legacy conventions, full_solver.py:574-610,856-881)
    for pd in ("x", "y"):
        lengths_p = {"x": (20e-3, 10e-3, 10e-3), "y": (10e-3, 20e-3, 10e-3)}[pd]
        xp, yp, zp = axes(lengths_p, ne.shape)
        dp = fs.ScalarDomain(xp, yp, zp, extent, probing_direction=pd)
        dp.external_ne(ne)
        dp.calc_dndr(lwl)
        np.random.seed(3)
        s0p = fs.init_beam(64, 4e-3, 5e-5, extent, "circular", pd)
        sfp = rk4(dp, s0p, 120)
        rfp, _ = fs.ray_to_Jonesvector(sfp, extent, probing_direction=pd)
        g3d.update({f"{pd}_x": xp, f"{pd}_y": yp, f"{pd}_z": zp, f"{pd}_s0": s0p, f"{pd}_sf": sfp, f"{pd}_rf": rfp})
    np.savez_compressed(os.path.join(OUT, "g3_turb.npz"), **g3d)

    # ---------------------------------------------------------------- G6: attenuation + Faraday channels (9-vector ODE)
    lengths, dims, extent = (10e-3, 10e-3, 20e-3), (24, 20, 28), 10e-3
    x, y, z = axes(lengths, dims)
    rng6 = np.random.default_rng(21)
    ne6 = gaussian_column(x, y, z, ne0=4e25, LR=3e-3) * (1 + 0.3 * np.cos(2 * np.pi * z / 7e-3))[None, None, :] + 1e24
    XX, YY, ZZ = np.meshgrid(x, y, z, indexing="ij")
    Te6 = 50.0 + 150.0 * np.exp(-(XX ** 2 + YY ** 2) / (4e-3) ** 2)            # eV
    Z6 = 3.0 + 2.0 * np.cos(2 * np.pi * ZZ / 9e-3) ** 2
    B6 = np.stack([5.0 * YY / 5e-3, -5.0 * XX / 5e-3, 10.0 * np.exp(-(XX ** 2 + YY ** 2) / (3e-3) ** 2)], axis=-1)   # T
    dom = fs.ScalarDomain(x, y, z, extent, B_on=True, inv_brems=True, phaseshift=True)
    dom.external_ne(ne6); dom.external_B(B6); dom.external_Te(Te6); dom.external_Z(Z6)
    dom.calc_dndr(lwl)
    dom.set_up_interps()
    s6 = probe_states(rng6, 2048, lengths)
    g6 = dict(x=x, y=y, z=z, extent=extent, lwl=lwl, ne=ne6, Te=Te6, Z=Z6, B=B6, s=s6, kappa=dom.kappa(),
              dsdt=fs.dsdt(0.0, s6.ravel().copy(), dom).reshape(9, -1))
    np.random.seed(6)
    s0 = fs.init_beam(128, 4e-3, 2e-3, extent, "circular", "z")
    g6["s0"], g6["rk4_nsteps"] = s0, 120
    g6["rk4_sf"] = rk4(dom, s0, 120)
    g6["rk4_rf"], g6["rk4_Jf"] = fs.ray_to_Jonesvector(g6["rk4_sf"], extent, probing_direction="z")
    # the shipped solver (joint RK45 over all 9N rows) and the per-ray variant, channels on
    rf6, Jf6 = quiet(dom.solve, s0.copy(), return_E=True)
    g6.update(joint_sf=dom.sf, joint_rf=rf6, joint_Jf=Jf6, joint_log=joint_log(fs, dom, s0))
    from scipy.integrate import solve_ivp as _ivp
    tt = np.linspace(0.0, np.sqrt(8.0) * extent / fs.c, 2)
    pr_sf, pr_nfev = np.empty((9, 16)), np.empty(16, dtype=np.int64)
    for i in range(16):
        sol = _ivp(lambda t, yy: fs.dsdt(t, yy, dom), [0, tt[-1]], s0[:, i].copy(), t_eval=tt)
        pr_sf[:, i], pr_nfev[i] = sol.y[:, -1], sol.nfev
    g6.update(perray_sf=pr_sf, perray_nfev=pr_nfev)
    np.savez_compressed(os.path.join(OUT, "g6_channels.npz"), **g6)

    # ---------------------------------------------------------------- G4: optics + detector
    rng = np.random.default_rng(5)
    n = 5000
    r0 = np.zeros((4, n))
    r0[0], r0[2] = rng.normal(0, 3e-3, n), rng.normal(0, 3e-3, n)          # metres
    r0[1], r0[3] = rng.normal(0, 2.5e-2, n), rng.normal(0, 2.5e-2, n)      # rad: many rays hit the stops
    r0[1, :1500] *= 0.02; r0[3, :1500] *= 0.02                             # ... and many pass DF/LF stops
    g4 = dict(r0=r0)
    def run(cls, meth, **kw):
        o = cls(r0.copy(), L=400, R=25)
        getattr(o, meth)(**kw)
        return o
    for tag, cls, meth, kw in [("shadow_single", rtm.Shadowgraphy, "single_lens_solve", {}),
                               ("shadow_two", rtm.Shadowgraphy, "two_lens_solve", {}),
                               ("schlieren_DF", rtm.Schlieren, "DF_solve", {"R": 1}),
                               ("schlieren_LF", rtm.Schlieren, "LF_solve", {"R": 1}),
                               ("refracto_incoherent", rtm.Refractometry, "incoherent_solve", {})]:
        o = run(cls, meth, **kw)
        g4[tag + "_rf"] = o.rf
        for bs in (25, 8):
            o.histogram(bin_scale=bs)
            g4[f"{tag}_H{bs}"] = o.H
    # element functions one by one (incl. knife edge and rect aperture AND-quirk)
    rmm = rtm.m_to_mm(r0[:, 1000:2000])
    g4["el_distance"] = rtm.distance(rmm.copy(), 123.0)
    g4["el_lens"] = rtm.lens(rmm.copy(), 200.0, 133.0)
    g4["el_circ_ap"] = rtm.circular_aperture(rmm.copy(), 4.0)
    g4["el_circ_stop"] = rtm.circular_stop(rmm.copy(), 4.0)
    g4["el_rect_ap"] = rtm.rect_aperture(rmm.copy(), 3.0, 2.0)
    g4["el_knife_y"] = rtm.knife_edge(rmm.copy(), 0.5, "y", 1)
    g4["el_knife_x"] = rtm.knife_edge(rmm.copy(), -0.5, "x", -1)
    # coherent chains: E from the G2 solve (phase accumulated), 384 rays, synthetic E on 2000 rays
    E = np.zeros((2, 2000), dtype=complex)
    ph = rng.random(2000) * 40
    E[1] = np.cos(ph) + 1j * np.sin(ph)
    E[0] = 0.1 * E[1] * np.exp(0.3j)
    rc = r0[:, :2000].copy(); rc[1] *= 0.02; rc[3] *= 0.02
    it = rtm.Interferometry(rc.copy(), E=E.copy(), L=400, R=25)
    it.two_lens_solve(wl=lwl)
    it.interferogram(bin_scale=40)
    g4.update(coh_r0=rc, coh_E=E, interf_rf=it.rf, interf_rE=it.rE, interf_H40=it.H)
    rfm = rtm.Refractometry(rc.copy(), E=E.copy(), L=400, R=25)
    rfm.coherent_solve(wl=lwl)
    g4.update(refr_coh_rf=rfm.rf, refr_coh_rE=rfm.rE)
    np.savez_compressed(os.path.join(OUT, "g4_optics.npz"), **g4)

    # ---------------------------------------------------------------- G5: docstring KATs + field fixture
    # SLAB test of full_solver.py:56-82 at reduced resolution (65^3) and NULL test (full_solver.py:12-54)
    N_V = 32; M_V = 2 * N_V + 1; ext = 5.0e-3
    a = np.linspace(-ext, ext, M_V)
    kat = {}
    for name in ("null", "slab"):
        dom = fs.ScalarDomain(a, a, a, ext)
        dom.test_null() if name == "null" else dom.test_slab(s=10, n_e0=1e25)
        dom.calc_dndr()                                       # default lwl 1053e-9
        np.random.seed(4)
        s0 = fs.init_beam(200, 5e-3, 0.5e-3, ext, "circular", "z")
        rf = quiet(dom.solve, s0.copy())
        kat[name + "_s0"], kat[name + "_rf"], kat[name + "_sf"] = s0, rf, dom.sf
    # field-generator fixture in the spirit of evaluation/sergio_testing/integratedPy.npy (reduced grid)
    xs, ys, zs = np.linspace(-5e-3, 5e-3, 20), np.linspace(-5e-3, 5e-3, 200), np.linspace(-5e-3, 5e-3, 20)
    dom = fs.ScalarDomain(xs, ys, zs, 5e-3)
    dom.test_linear_cos(s1=-1, s2=1, n_e0=1e26, Ly=5e-3)
    kat["linear_cos_integrated"] = dom.ne.sum(axis=2)
    np.savez_compressed(os.path.join(OUT, "g5_kat.npz"), axis=a, extent=ext, **kat)

    minimal_fixture()
    reference_fixture()
    fresnel_fixture()
    diagnostics_fixture()
    beam_fixture()
    domain_fixture()
    propagator_fixture()
    current_solve_fixture()
    louis_fixture()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)) // 1024, "KiB")


if __name__ == "__main__":
    if sys.argv[1:] == ["fresnel"]:
        fresnel_fixture()
    elif sys.argv[1:] == ["diagnostics"]:
        diagnostics_fixture()
        beam_fixture()
        domain_fixture()
        propagator_fixture()
        current_solve_fixture()
    elif sys.argv[1:] == ["louis"]:
        louis_fixture()
    elif sys.argv[1:] == ["minimal"]:
        minimal_fixture()
        reference_fixture()
    else:
        main()
