"""CPU self-test: the product's per-ray math (csrc/ray_core.h, field_prep.h), compiled for the host by
tests/host_harness.cpp, against the golden vectors of the real reference.  The -m gpu tests repeat these
through the C ABI on the device; this file is what can run in a container without a GPU."""
import numpy as np
import pytest

from conftest import rel_err
from harness import Harness
from oracle import synthpy_oracle as O

C_LIGHT = 299792458.0
omega_of = lambda lwl: 2 * np.pi * (C_LIGHT / lwl)


@pytest.fixture(scope="module")
def H():
    return Harness()


def test_field_stencil_bit_equal_numpy(H, golden):
    g = golden("g1_rhs")
    for march in (0, 1, 2):
        f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(float(g["lwl"])), march_axis=march, phase=True)
        gx, gy, gz, aux = f.export()
        assert np.array_equal(gx, g["gradx"]) and np.array_equal(gy, g["grady"]) and np.array_equal(gz, g["gradz"])
    # uniform-spacing branch of np.gradient: axes whose float32 spacings are all equal
    x = np.arange(12) * 0.25 - 1.0
    ne = 1e24 * np.random.default_rng(0).random((12, 12, 12))
    d = O.Domain(x, x, x, 1.0)
    d.external_ne(ne)
    d.calc_dndr(1064e-9)
    f = H.field(ne, x, x, x, omega_of(1064e-9))
    gx, gy, gz, _ = f.export()
    assert np.array_equal(gx, d.grads[0]) and np.array_equal(gy, d.grads[1]) and np.array_equal(gz, d.grads[2])


def test_rhs_matches_reference(H, golden):
    g = golden("g1_rhs")
    for ph in (False, True):
        f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(float(g["lwl"])), phase=ph, f64=ph)
        out = f.rhs(g["s"])
        ref = g["dsdt_phase%d" % ph]
        assert np.array_equal(out[:3], ref[:3])
        assert np.array_equal(out[3:6] == 0, ref[3:6] == 0)          # identical in/out-of-bounds decisions
        assert rel_err(out[3:6], ref[3:6], floor=1e3) < 1e-11      # polynomial form of the trilinear: few-ulp of the corner scale
        if ph:
            # the reference interpolates n ~ 1 and subtracts 1 afterwards: its own rounding floor is a few
            # eps * omega in absolute terms (we interpolate n-1, which is more accurate at low density)
            assert np.all(np.abs(out[7] - ref[7]) <= 1e-15 * f.omega + 1e-12 * np.abs(ref[7]))
    # float32 aux lane (stored AND interpolated in float32): ~1e-7 of the largest |n - 1| among the cell's corners
    # (documented fast mode; the accumulated phase is then good to ~1e-7 relative), still exactly zero outside
    f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(float(g["lwl"])), phase=True, f64=False)
    out = f.rhs(g["s"])
    ref7 = g["dsdt_phase1"][7]
    nm1 = np.abs(np.sqrt(1.0 - (5.64e4 * np.sqrt(g["ne"] * 1e-6) / f.omega) ** 2) - 1.0)
    ax = [np.float64(np.float32(g[k])) for k in ("x", "y", "z")]
    idx = [np.clip(np.searchsorted(a, g["s"][k], side="right") - 1, 0, len(a) - 2) for k, a in enumerate(ax)]
    cell_max = np.zeros(g["s"].shape[1])
    for du in (0, 1):
        for dv in (0, 1):
            for dw in (0, 1):
                cell_max = np.maximum(cell_max, nm1[idx[0] + du, idx[1] + dv, idx[2] + dw])
    assert np.all(np.abs(out[7] - ref7) <= 1e-15 * f.omega + 3e-7 * f.omega * cell_max)
    inside = np.all([(g["s"][k] >= a[0]) & (g["s"][k] <= a[-1]) for k, a in enumerate(ax)], axis=0)
    assert np.all(out[7][~inside] == 0)


def test_non_uniform_axes(H):
    """The legacy API takes arbitrary ascending axes (full_solver.py:102-120): the cell search walks the real table."""
    rng = np.random.default_rng(8)
    x = np.cumsum(rng.uniform(0.5, 2.0, 19)); x = (x - x.mean()) * 1e-3 / 3
    y = np.sort(rng.uniform(-4e-3, 4e-3, 15)); z = np.linspace(-1, 1, 23) ** 3 * 8e-3
    ne = 1e25 * (1 + 0.5 * rng.random((19, 15, 23)))
    d = O.Domain(x, y, z, 8e-3)
    d.external_ne(ne)
    d.calc_dndr(1064e-9)
    f = H.field(ne, x, y, z, omega_of(1064e-9))
    gx, gy, gz, _ = f.export()
    assert np.array_equal(gx, d.grads[0]) and np.array_equal(gy, d.grads[1]) and np.array_equal(gz, d.grads[2])
    s = np.zeros((9, 3000))
    s[0], s[1], s[2] = rng.uniform(x[0] * 1.1, x[-1] * 1.1, 3000), rng.uniform(-4.4e-3, 4.4e-3, 3000), rng.uniform(-8.5e-3, 8.5e-3, 3000)
    s[0, :19], s[2, 19:42] = np.float64(np.float32(x)), np.float64(np.float32(z))         # exactly on nodes
    s[3:6] = 1e8
    ref = d.dsdt(0.0, s.ravel().copy()).reshape(9, -1)
    out = f.rhs(s)
    assert np.array_equal(out[3:6] == 0, ref[3:6] == 0)
    assert np.max(np.abs(out[3:6] - ref[3:6])) < 1e-11 * np.abs(ref[3:6]).max()


def test_cell_cache_walk_equals_fresh_lookups(H):
    """A cell cache dragged through an arbitrary walk -- single-cell moves in both directions on every axis, jumps,
    excursions outside the grid and back, points exactly on nodes -- returns what a fresh lookup returns (bit for bit
    away from nodes), for the neighbour relocation of the fixed-step kernel and for the direct search of the adaptive ones."""
    rng = np.random.default_rng(21)
    x = np.cumsum(rng.uniform(0.5, 2.0, 17)); x = (x - x.mean()) * 1e-3 / 3
    y = np.linspace(-3e-3, 3e-3, 13); z = np.linspace(-1, 1, 29) ** 3 * 6e-3
    ne = 1e25 * (1 + 0.5 * rng.random((17, 13, 29)))
    f = H.field(ne, x, y, z, omega_of(1064e-9))
    ax = [np.float64(np.float32(a)) for a in (x, y, z)]
    n = 40000
    pts = np.zeros((3, n))
    p = np.array([a[len(a) // 2] for a in ax]) + 1e-7
    for i in range(n):
        u = rng.random()
        if u < 0.70:                                   # small move: stays or crosses one face, either direction
            k = rng.integers(3)
            p[k] += rng.normal() * 0.6 * np.median(np.diff(ax[k]))
        elif u < 0.85:                                 # jump anywhere, sometimes outside
            p = np.array([rng.uniform(a[0] - 0.2 * (a[-1] - a[0]), a[-1] + 0.2 * (a[-1] - a[0])) for a in ax])
        elif u < 0.95:                                 # exactly on a node of one axis
            k = rng.integers(3)
            p[k] = ax[k][rng.integers(len(ax[k]))]
        else:                                          # one ulp either side of a node
            k = rng.integers(3)
            p[k] = np.nextafter(ax[k][rng.integers(len(ax[k]))], rng.choice([-np.inf, np.inf]))
        pts[:, i] = p
    s = np.zeros((9, n)); s[0:3] = pts; s[3:6] = 1.0
    fresh = f.rhs(s)
    assert 0.05 < np.mean(np.all(fresh[3:6] == 0, axis=0)) < 0.6          # the walk spends real time outside the grid too
    scale = np.abs(fresh[3:6]).max()
    for near in (True, False):
        got = f.rhs_walk(s, near)
        assert np.array_equal(got[3:6] == 0, fresh[3:6] == 0), near           # same in / out of grid decisions
        differs = np.any(got[3:6] != fresh[3:6], axis=0)
        # a point ON a node may be served by the cell below it with weight 1 - ulp instead of the cell above with
        # weight 0 (the interpolant is continuous there): equal to rounding, and only there
        assert differs.mean() < 0.01 and np.max(np.abs(got[3:6] - fresh[3:6])) < 1e-14 * scale, near
        on_node = np.zeros(n, dtype=bool)
        for k in range(3):
            d = np.abs(pts[k][:, None] - ax[k][None, :]).min(axis=1)
            on_node |= d <= 4 * np.spacing(np.abs(pts[k]))
        assert not np.any(differs & ~on_node), near


def test_rk4_matches_reference_rhs_loop(H, golden):
    for name, ph in (("g2_expcos", True), ("g3_turb", False)):
        g = golden(name)
        ext, n = float(g["extent"]), int(g["rk4_nsteps"])
        f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(float(g["lwl"])), phase=ph, f64=ph)
        h = np.sqrt(8.0) * ext / C_LIGHT / n
        sf, steps = f.rk4(g["s0"][:, :128], n, h)
        assert (steps == n).all()
        assert rel_err(sf[:6], g["rk4_sf"][:6], floor=1e-6) < 1e-11
        rf = f.exit(sf, 2, 0, 1, ext)
        ref_rf, _ = O.ray_to_jones(g["rk4_sf"], ext)
        assert rel_err(rf, ref_rf, floor=1e-7) < 1e-9
        if ph:
            assert rel_err(sf[7], g["rk4_sf"][7], floor=1e-3) < 1e-11
        # early exit leaves every ray on the same straight line
        sfe, st = f.rk4(g["s0"][:, :128], n, h, early=True)
        assert st.max() < n
        assert rel_err(f.exit(sfe, 2, 0, 1, ext), ref_rf, floor=1e-7) < 1e-9


def test_rk4_other_probing_directions(H, golden):
    g = golden("g3_turb")
    ext = float(g["extent"])
    for pd, (p, a, b) in {"x": (0, 1, 2), "y": (1, 0, 2)}.items():
        f = H.field(g["ne"], g[pd + "_x"], g[pd + "_y"], g[pd + "_z"], omega_of(float(g["lwl"])), march_axis=p)
        h = np.sqrt(8.0) * ext / C_LIGHT / 120
        sf, _ = f.rk4(g[pd + "_s0"], 120, h)
        assert rel_err(sf[:6], g[pd + "_sf"][:6], floor=1e-6) < 1e-11
        assert rel_err(f.exit(sf, p, a, b, ext), g[pd + "_rf"], floor=1e-7) < 1e-9


def test_rk45_per_ray_matches_solve_ivp(H, golden):
    g = golden("g2_expcos")
    ext = float(g["extent"])
    f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(float(g["lwl"])), phase=True, f64=True)
    t_end = np.sqrt(8.0) * ext / C_LIGHT
    # (a) SciPy's default tolerances (what the reference ships): identical accept/reject sequences
    for tag, (rtol, atol) in {"def": (1e-3, 1e-6)}.items():
        sf, att, nfev = f.rk45(g["s0"][:, :32], t_end, rtol, atol)
        assert np.array_equal(nfev, g["perray_nfev_" + tag])
        ref = g["perray_sf_" + tag]
        # Same algorithm, same step sequence.  Bitwise agreement is not attainable for an ADAPTIVE method: the
        # controller scales h by err_norm**-0.2 and err_norm is a cancelling sum (5th minus 4th order) whose
        # rounding noise (order of BLAS summation inside np.dot) reaches 1e-8 relative when err_norm ~ 1e-4,
        # i.e. h differs by ~4e-9 relative and the (rtol 1e-3) truncation error moves with it.  Hence the bar is
        # 1e-9 of the natural scales: domain size, c, and the phase magnitude.
        assert np.max(np.abs(sf[:3] - ref[:3])) < 1e-9 * ext
        assert np.max(np.abs(sf[3:6] - ref[3:6])) < 1e-9 * C_LIGHT
        assert np.max(np.abs(sf[7] - ref[7])) < 1e-9 * np.abs(ref[7]).max()
        rf, rf_ref = f.exit(sf, 2, 0, 1, ext), O.ray_to_jones(ref, ext)[0]
        assert np.max(np.abs(rf - rf_ref)) < 1e-9
    # (b) tight tolerances (1e-7 / 1e-9, the reference's "intended" diffrax values).  In the free-flight legs the
    # true local error is zero, so SciPy's err_norm is pure rounding noise (~1e-5) and its step factor
    # 0.9 * noise**-0.2 ~ 9.2 is noise too: the reference's own step sequence is not reproducible across BLAS
    # builds.  Parity there means: same solution to within the requested tolerance, and no worse than SciPy
    # against a near-converged (rtol 1e-11) solve of the same RHS.
    sf, att, nfev = f.rk45(g["s0"][:, :32], t_end, 1e-7, 1e-9)
    ref = g["perray_sf_tight"]
    assert abs(int(nfev.sum()) - int(g["perray_nfev_tight"].sum())) < 0.05 * g["perray_nfev_tight"].sum()
    rf, rf_ref = f.exit(sf, 2, 0, 1, ext), O.ray_to_jones(ref, ext)[0]
    assert np.max(np.abs(rf[[0, 2]] - rf_ref[[0, 2]])) < 1e-7 * ext
    assert np.max(np.abs(rf[[1, 3]] - rf_ref[[1, 3]])) < 5e-7
    rf_conv = O.ray_to_jones(g["perray_sf_conv"], ext)[0]
    err_ours = np.abs(rf[:, :8] - rf_conv).max(axis=1)
    err_scipy = np.abs(rf_ref[:, :8] - rf_conv).max(axis=1)
    # (both sequences are noise-driven there, so either error is a draw around the tolerance level: factor 3, not 1)
    assert np.all(err_ours <= 3 * err_scipy + 1e-12)


def test_attenuation_and_faraday_channels(H, golden):
    g = golden("g6_channels")
    lwl, ext = float(g["lwl"]), float(g["extent"])
    f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(lwl), phase=True, f64=True)
    f.attach(g["kappa"], g["ne"], g["B"])
    verdet = 2.62e-13 * lwl ** 2
    out = f.rhs_ext(g["s"], verdet)
    ref = g["dsdt"]
    assert np.array_equal(out[6] == 0, ref[6] == 0) and np.array_equal(out[8] == 0, ref[8] == 0)
    assert np.max(np.abs(out[6] - ref[6])) < 1e-12 * np.abs(ref[6]).max()
    assert np.max(np.abs(out[8] - ref[8])) < 1e-12 * np.abs(ref[8]).max()
    n = int(g["rk4_nsteps"])
    sf = f.rk4_ext(g["s0"], n, np.sqrt(8.0) * ext / C_LIGHT / n, verdet)
    assert rel_err(sf[:6], g["rk4_sf"][:6], floor=1e-6) < 1e-10
    assert np.max(np.abs(sf[6] - g["rk4_sf"][6])) < 1e-10 * np.abs(g["rk4_sf"][6]).max()       # amplitude
    assert np.max(np.abs(sf[7] - g["rk4_sf"][7])) < 1e-10 * np.abs(g["rk4_sf"][7]).max()       # phase
    assert np.max(np.abs(sf[8] - g["rk4_sf"][8])) < 1e-10 * np.abs(g["rk4_sf"][8]).max()       # polarisation
    assert np.ptp(g["rk4_sf"][6]) > 1e-3 and np.abs(g["rk4_sf"][8]).max() > 1e-6                # channels are live
    # adaptive, per ray, all nine rows in the error norm: same accept/reject sequence as solve_ivp
    sf, nfev = f.rk45_ext(g["s0"][:, :16], np.sqrt(8.0) * ext / C_LIGHT, verdet)
    assert np.array_equal(nfev, g["perray_nfev"])
    ref = g["perray_sf"]
    assert np.max(np.abs(sf[:3] - ref[:3])) < 1e-9 * ext and np.max(np.abs(sf[3:6] - ref[3:6])) < 1e-9 * C_LIGHT
    for row in (6, 7, 8):
        assert np.max(np.abs(sf[row] - ref[row])) < 1e-9 * np.abs(ref[row]).max()


def test_fp32_mode_within_1e4(H, golden):
    """FP32 state/arithmetic (the JAX generation's default precision) against the FP64 oracle.  The golden rays
    start at z = -extent, which is OUTSIDE the float32-rounded z[0] in float64 but ON it once rounded to
    float32 -- a knife-edge the two precisions legitimately resolve differently -- so the rays are first moved
    back along their (straight) path by a fraction of a cell."""
    g = golden("g3_turb")
    ext, n = float(g["extent"]), int(g["rk4_nsteps"])
    s0 = g["s0"][:, :128].copy()
    s0[:3] -= s0[3:6] * (2e-4 / C_LIGHT)
    d = O.Domain(g["x"], g["y"], g["z"], ext)
    d.external_ne(g["ne"])
    d.calc_dndr(float(g["lwl"]))
    h = np.sqrt(8.0) * ext / C_LIGHT / n
    ref_rf, _ = O.ray_to_jones(d.solve_rk4(s0, n, h)[0], ext)
    f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(float(g["lwl"])))
    sf, _ = f.rk4(s0, n, h, fp32=True)
    rf = f.exit(sf, 2, 0, 1, ext)
    assert np.max(np.abs(rf[[0, 2]] - ref_rf[[0, 2]])) < 1e-4 * np.abs(ref_rf[[0, 2]]).max()
    assert np.max(np.abs(rf[[1, 3]] - ref_rf[[1, 3]])) < 1e-4 * np.abs(ref_rf[[1, 3]]).max()


def test_optics_elements_and_chains(H, golden):
    g = golden("g4_optics")
    rmm = O.m_to_mm(g["r0"][:, 1000:2000])
    for key, op in [("el_distance", ("travel", 123.0)), ("el_lens", ("lens", 200.0, 133.0)),
                    ("el_circ_ap", ("circ_ap", 4.0)), ("el_circ_stop", ("circ_stop", 4.0)),
                    ("el_rect_ap", ("rect_ap", 3.0, 2.0)), ("el_knife_y", ("knife", 0.5, 2, 1)),
                    ("el_knife_x", ("knife", -0.5, 0, -1))]:
        out = H.optics(rmm, [op], input_mm=True)
        assert rel_err(out, g[key], floor=1e-9) < 1e-12, key      # np.matmul (BLAS) fuses multiply-adds
    for tag in ("shadow_single", "shadow_two", "schlieren_DF", "schlieren_LF", "refracto_incoherent"):
        out = H.optics(g["r0"], O.chain(tag))
        assert rel_err(out, g[tag + "_rf"], floor=1e-3) < 1e-11, tag      # imaging chains cancel: abs 1e-14 mm
    lwl = 1064e-9
    for tag, key in (("interf_two", "interf"), ("refracto_coherent", "refr_coh")):
        r, E = H.optics(g["coh_r0"], O.chain(tag), jf=g["coh_E"], wavelength=lwl)
        assert rel_err(r, g[key + "_rf"], floor=1e-3) < 1e-11
        assert np.array_equal(np.isnan(E.real), np.isnan(g[key + "_rE"].real))
        m = ~np.isnan(E.real)
        assert np.max(np.abs(E[m] - g[key + "_rE"][m])) < 1e-6            # k*r ~ 1e8 rad: 1 ulp of the argument ~ 1e-8



def _host_engine(H, monkeypatch):
    """Route the diagnostics classes' three engine calls to the host build of the SAME optics code (csrc/ray_core.h::
    run_optics through tests/host_harness.cpp), so that the Python composition above the C ABI is testable without a GPU."""
    import torch
    from synthpy_b200 import engine

    def to_device(a, dtype=torch.float64):
        return torch.as_tensor(np.ascontiguousarray(a.numpy() if isinstance(a, torch.Tensor) else a)).to(dtype)

    def optics_image(rf, ops, *, jf=None, image=None, wavelength=0.0, input_mm=False, want_rays=True):
        assert image is None
        out = H.optics(rf.numpy(), ops, jf=None if jf is None else jf.numpy(), wavelength=wavelength or 0.0, input_mm=input_mm)
        if jf is None:
            return torch.from_numpy(out), None
        return torch.from_numpy(out[0]), torch.from_numpy(out[1])

    monkeypatch.setattr(engine, "to_device", to_device)
    monkeypatch.setattr(engine, "optics_image", optics_image)
    monkeypatch.setattr(engine, "require_cuda", lambda: None)


def test_current_generation_diagnostics_host(H, golden, monkeypatch):
    """g9 (src/simulator/diagnostics.py run from its own source): the product's optics math on every layout of the current
    API, the reference beam, and the two-pass composition behind ``Refractometry.coherent_solve(generation='current')``."""
    from synthpy_b200 import diagnostics as D
    g = golden("g9_diagnostics")
    rf, Jf, lwl = g["rf"], g["Jf"], float(g["lwl"])
    kw = dict(L=float(g["L"]), R=float(g["R"]), focal_plane=float(g["focal_plane"]))

    def same_field(E, ref):
        assert np.array_equal(np.isnan(E.real), np.isnan(ref.real))
        m = ~np.isnan(ref.real)
        assert m.sum() > 100 and np.max(np.abs(E[m] - ref[m])) < 1e-6         # k * path ~ 1e9 rad: an ulp of the argument ~ 1e-7

    for meth, name in (("single_lens_solve", "shadow_single"), ("two_lens_solve", "shadow_two"), ("DF_solve", "schlieren_DF"),
                       ("LF_solve", "schlieren_LF"), ("incoherent_solve", "refracto_incoherent")):
        assert rel_err(H.optics(rf, D.chain_ops(name, **kw)), g[meth + "_rf"], floor=1e-3) < 1e-11, meth
    _, E = H.optics(rf, [("ref_beam", 7, 60)], jf=Jf)
    assert rel_err(E.view(np.float64), g["ref_beam_7_60_Jf"].view(np.float64), floor=1e-3) < 1e-12
    r, E = H.optics(rf, [("ref_beam", 10, 20)] + D.chain_ops("interf_two", **kw), jf=Jf, wavelength=lwl)
    assert rel_err(r, g["interf_rf"], floor=1e-3) < 1e-11
    same_field(E, g["interf_Jf"])

    _host_engine(H, monkeypatch)
    for tag, k in (("coherent_solve", kw), ("coherent_R6", dict(L=300, R=6, focal_plane=0))):
        d = D.Refractometry(lwl, rf.copy(), Jf.copy(), **k)
        d.coherent_solve()
        assert rel_err(d.rf, g[tag + "_rf"], floor=1e-3) < 1e-11, tag
        same_field(d.Jf, g[tag + "_Jf"])
    d = D.Refractometry(lwl, rf.copy(), Jf.copy(), **kw)                     # the legacy layout is still there, and different
    d.coherent_solve(generation="legacy")
    ro, Eo = O.run_chain(rf, O.chain("refracto_coherent", **kw), E=Jf, wl=lwl)
    assert rel_err(d.rf, ro, floor=1e-3) < 1e-11 and np.nanmax(np.abs(d.rf - g["coherent_solve_rf"])) > 1.0
    with pytest.raises(ValueError):
        d.coherent_solve(generation="jax")
    with pytest.raises(ValueError):
        D.Refractometry(lwl, rf.copy(), **kw).coherent_solve()
    it = D.Interferometry(lwl, rf.copy(), Jf.copy(), **kw)
    it.two_lens_solve()
    assert rel_err(it.rf, g["interf_rf"], floor=1e-3) < 1e-11
    same_field(it.Jf, g["interf_Jf"])



def test_current_generation_propagator_functions(H, golden):
    """g12: src/simulator/propagator.py's array-level functions executed from their own source (oracle/gen_golden.py::
    import_simulator): the RHS as the current generation evaluates it (gradient of ne / (3.142e-4 omega^2) in float64 at every
    call, its own interpolator), ``ray_to_Jonesvector`` with the current output-axis convention and ``back_propogate``."""
    from synthpy_b200 import propagator as P
    g, g1 = golden("g12_propagator"), golden("g1_rhs")
    assert abs(float(g["omega"]) / omega_of(float(g1["lwl"])) - 1) < 1e-12
    # RHS: same velocities, same in/out-of-grid decisions; accelerations equal to the float32 rounding of the gradient table
    # that the legacy generation (and the packed field here) keeps -- 1.7e-7 of the largest one, the same distance as
    # between the two upstream generations themselves
    f = H.field(g1["ne"], g1["x"], g1["y"], g1["z"], float(g["omega"]))
    out, ref = f.rhs(g1["s"]), g["dsdt"]
    assert np.array_equal(out[:3], ref[:3]) and np.array_equal(out[3:6] == 0, ref[3:6] == 0) and not ref[6:].any()
    scale = np.abs(ref[3:6]).max()
    assert np.abs(out[3:6] - ref[3:6]).max() < 3e-7 * scale
    assert np.abs(g1["dsdt_phase0"][3:6] - ref[3:6]).max() < 3e-7 * scale
    # NRL inverse-bremsstrahlung rate (propagator.py:30-61), densities up to above critical: the host field preparation is identical
    from synthpy_b200 import engine
    assert np.array_equal(engine.kappa_grid(g["k_ne"], g["k_Te"], g["k_Z"], float(g["omega"])), g["kappa"]) and (g["kappa"] > 0).all()
    with np.errstate(invalid="ignore"):
        assert np.array_equal(np.sqrt(1.0 - (5.64e4 * np.sqrt(g["k_ne"] * 1e-6) / float(g["omega"])) ** 2), g["n_refrac"], equal_nan=True)
    # exit plane, the product's projection code with the axis maps the Python layer passes
    for pd, p in (("x", 0), ("y", 1), ("z", 2)):
        st = g["sf_" + pd]
        a, b = P._out_axes(pd, "current")
        rf = f.exit(st, p, a, b, 5e-3)
        assert rel_err(rf, g["rtj_%s_0_p" % pd], floor=1e-7) < 1e-13, pd
        rj = O.ray_to_jones(st, 5e-3)[1]                                    # the field does not depend on the direction
        assert np.max(np.abs(rj - g["rtj_%s_0_J" % pd])) < 1e-12 and np.array_equal(g["rtj_%s_0_J" % pd], g["rtj_%s_1_J" % pd])
        keep = g["rtj_%s_1_p" % pd]
        assert np.array_equal(keep[0], st[a]) and np.array_equal(keep[2], st[b]) and np.array_equal(keep[[1, 3]], g["rtj_%s_0_p" % pd][[1, 3]])
        # back_propogate: rows 0..2 hold (a, b) on the exit plane; upstream stores them in ray_p order for 'y' (z, plane, x)
        bp = g["bp_" + pd]
        order = {"x": (1, 2), "y": (0, 2), "z": (0, 1)}[pd]
        assert np.all(bp[p] == 5e-3) and np.array_equal(bp[3:], st[3:])
        assert rel_err(bp[order[0]], g["rtj_%s_0_p" % pd][0], floor=1e-7) < 1e-13 and rel_err(bp[order[1]], g["rtj_%s_0_p" % pd][2], floor=1e-7) < 1e-13



class _HostSlabBackend:
    """``out_of_core.trace_slabs`` on the host build of the ray code (tests/host_harness.cpp): same planner and driver as
    the device path, the arithmetic of csrc/ray_core.h + field_prep.h compiled for the CPU."""

    def __init__(self, H, omega, p, out_axes, extent, phase=False):
        self.H, self.omega, self.p, self.out_axes, self.extent, self.phase = H, omega, p, out_axes, extent, phase
        self.launches, self.vscale = [], 1.0

    def to_state(self, s0):
        return np.array(s0, dtype=np.float64)

    def live_range(self, s, lo, hi):
        pos, vel = s[:3], s[3:6]
        lo, hi = np.array(lo)[:, None], np.array(hi)[:, None]
        gone = (((pos > hi) & (vel >= 0)) | ((pos < lo) & (vel <= 0))).any(0)
        live = ~gone & np.isfinite(s[:6]).all(0)
        if not live.any():
            return None
        z, vz = pos[self.p][live], vel[self.p][live]
        return float(z.min()), float(z.max()), float(np.abs(vz).max()) * self.vscale

    def field(self, ne_slab, axes):
        return self.H.field(np.ascontiguousarray(ne_slab), axes[0], axes[1], axes[2], self.omega, march_axis=self.p,
                            phase=self.phase, f64=self.phase)

    def steps(self, field, s, m, h, last, want_jf, channels):
        self.launches.append(m)
        sf, st = field.rk4(s, m, h, early=True)
        rf = field.exit(sf, self.p, self.out_axes[0], self.out_axes[1], self.extent) if last else None
        return sf, st.astype(np.int64), rf, None, None

    def exit(self, s, want_jf):
        raise AssertionError("not reached in these tests")

    def release(self, field):
        pass


def test_out_of_core_planner_and_driver(H, golden):
    """Slab-wise tracing (synthpy_b200/out_of_core.py) gives the one-region result BIT FOR BIT: states, steps per ray, exit
    rays -- turbulent field, probing along z, x and y, with the phase lane, slabs of 9 .. 14 of 32 planes."""
    from synthpy_b200 import out_of_core as OC
    g = golden("g3_turb")
    omega, ext = omega_of(float(g["lwl"])), float(g["extent"])
    for pd, p, pre, out_axes, planes, phase in (("z", 2, "", (0, 1), 9, False), ("z", 2, "", (0, 1), 14, True), ("x", 0, "x_", (1, 2), 11, False),
                                                  ("y", 1, "y_", (2, 0), 12, True)):
        axes = [g[pre + k] for k in "xyz"]
        s0 = g[pre + "s0"]
        cell = np.diff(axes[p]).min()
        h = 0.5 * cell / C_LIGHT
        n = int(np.ceil(np.sqrt(8.0) * ext / C_LIGHT / h))
        whole = H.field(g["ne"], *axes, omega, march_axis=p, phase=phase, f64=phase)
        sf, steps = whole.rk4(s0, n, h, early=True)
        rf = whole.exit(sf, p, out_axes[0], out_axes[1], ext)
        be = _HostSlabBackend(H, omega, p, out_axes, ext, phase)
        s_sl, st_sl, rf_sl, _, log, _ = OC.trace_slabs(be, s0, OC.array_source(g["ne"], pd), axes, pd, n, h, planes)
        assert len(log) >= 3 and log[0]["planes"][0] == 0 and log[-1]["planes"][1] == len(axes[p]) - 1, log
        assert all(b - a + 1 <= planes + 1 for a, b in (e["planes"] for e in log))
        assert np.array_equal(s_sl, sf, equal_nan=True) and np.array_equal(st_sl, steps) and np.array_equal(rf_sl, rf, equal_nan=True), pd
        assert steps.max() < n and sum(e["steps"] for e in log) == n              # early exit at work; every step accounted for
        if phase:
            assert np.abs(sf[7]).max() > 1.0
        assert len(be.launches) == len(log)                                        # no slab had to be redone
    # a planner that under-estimates the speed over-plans the steps: the check on the actual positions catches it, the slab is
    # redone with fewer steps, and the result is still the one-region result
    be = _HostSlabBackend(H, omega, p, out_axes, ext, phase)
    be.vscale = 0.45
    s_sl, st_sl, rf_sl, _, log, _ = OC.trace_slabs(be, s0, OC.array_source(g["ne"], pd), axes, pd, n, h, planes)
    assert len(be.launches) > len(log)
    assert np.array_equal(s_sl, sf, equal_nan=True) and np.array_equal(st_sl, steps) and np.array_equal(rf_sl, rf, equal_nan=True)
    # the same trace with the next slab read on a worker thread: identical result, every slab after the first served from the
    # prefetched window; a request outside the window falls through to the wrapped source; errors surface on the caller
    be = _HostSlabBackend(H, omega, p, out_axes, ext, phase)
    calls = []

    def slow_source(k0, k1, inner=OC.array_source(g["ne"], pd)):
        calls.append((k0, k1))
        return inner(k0, k1)
    pf = OC.PrefetchingSource(slow_source, len(axes[p]), p)
    s_pf, st_pf, rf_pf, _, log_pf, _ = OC.trace_slabs(be, s0, pf, axes, pd, n, h, planes)
    pf.close()
    assert np.array_equal(s_pf, sf, equal_nan=True) and np.array_equal(st_pf, steps) and np.array_equal(rf_pf, rf, equal_nan=True)
    assert pf.misses == 1 and pf.hits == len(log_pf) - 1 >= 2
    pf = OC.PrefetchingSource(slow_source, len(axes[p]), p, back=4)
    idx = [slice(None)] * 3
    idx[p] = slice(20, 26)
    assert np.array_equal(pf(0, 8), slow_source(0, 8)) and np.array_equal(pf(20, 26), g["ne"][tuple(idx)]) and pf.misses == 2
    pf.close()

    def broken(k0, k1):
        if k0 > 0:
            raise OSError("disk")
        return slow_source(k0, k1)
    pf = OC.PrefetchingSource(broken, len(axes[p]), p, back=4)
    pf(0, 8)
    with pytest.raises(OSError):
        pf(6, 12)
    # the planner alone
    zc = np.float64(np.float32(np.linspace(-1e-2, 1e-2, 64)))
    a, b, m, z_stop = OC.plan_slab(zc, -1e-2, -1e-2, C_LIGHT, 1e-13, 500, 16)
    assert (a, b) == (0, 15) and z_stop == zc[13] and m == int((zc[13] + 1e-2) / (1e-13 * C_LIGHT * 1.02))
    a, b, m, z_stop = OC.plan_slab(zc, zc[40] + 1e-6, zc[41], C_LIGHT, 1e-13, 500, 16)
    assert (a, b) == (38, 53) and zc[a + 2] <= zc[40] + 1e-6
    a, b, m, z_stop = OC.plan_slab(zc, zc[55], zc[56], C_LIGHT, 1e-13, 77, 16)
    assert b == 63 and m == 77 and z_stop == np.inf
    with pytest.raises(ValueError):
        OC.plan_slab(zc, zc[10], zc[30], C_LIGHT, 1e-13, 500, 16)                  # rays spread over more planes than a slab holds
    # an axis whose float32 spacings are NOT all equal while those of a short window are: the window grows until its stencil
    # choice (np.gradient's uniform / non-uniform formula) is the full axis's
    zz = np.concatenate([np.arange(40) * 2.0 ** -10, 40 * 2.0 ** -10 + np.arange(1, 25) * 2.0 ** -10 * (1 + 2.0 ** -20)])
    assert not OC.axis_is_uniform(zz) and OC.axis_is_uniform(zz[:16])
    a, b, m, _ = OC.plan_slab(zz, 0.0, 0.0, C_LIGHT, 1e-13, 500, 16)
    assert a == 0 and b == 40 and not OC.axis_is_uniform(zz[a:b + 1]) and OC.axis_is_uniform(zz[a:b])



def test_louis_layouts(H, golden):
    """g13: src/solvers-legacy/rtm_solver-louis.py run as it is (sympy-lambdified composite matrices).  Its four trains as
    op lists (``diagnostics.chain_ops('louis_*')``), the knife edge in the Fourier plane included."""
    from synthpy_b200 import diagnostics as D
    g = golden("g13_louis")
    kw = dict(L=float(g["L"]), R=float(g["R"]))
    for tag, extra in (("refractometer", {}), ("shadowgraphy", {"displacement": float(g["displacement"])}), ("schlieren", {})):
        r = H.optics(g["r0"], D.chain_ops("louis_" + tag, **kw, **extra), input_mm=True)
        assert rel_err(r, g[tag + "_rf"], floor=1e-3) < 1e-11, tag
        nx, ny = 3448 // 24, 2574 // 24
        ix, iy = H.bins(r[0], -9.0, 9.0, nx, True), H.bins(r[2], -6.75, 6.75, ny, True)
        ok = (ix >= 0) & (iy >= 0)
        Hh = np.zeros((ny, nx))
        np.add.at(Hh, (iy[ok], ix[ok]), 1.0)
        assert np.array_equal(Hh, g[tag + "_H"]) and Hh.sum() > 500, tag
    assert np.isnan(g["schlieren_rf"][0]).sum() > np.isnan(g["shadowgraphy_rf"][0]).sum() + 1000      # the knife edge cuts
    r, E = H.optics(g["r0"], D.chain_ops("louis_interferometer", **kw), jf=g["E"], wavelength=float(g["wl"]), input_mm=True)
    assert rel_err(r, g["interferometer_rf"], floor=1e-3) < 1e-11 and np.max(np.abs(E - g["interferometer_rE"])) < 1e-5


def test_bin_search_matches_numpy(H):
    rng = np.random.default_rng(3)
    for lo, hi, nb in [(-9.0, 9.0, 3448), (-6.75, 6.75, 2574), (-9.0, 9.0, 137), (-7.0, 6.0, 63)]:
        edges = np.linspace(lo, hi, nb + 1)
        v = np.concatenate([rng.uniform(lo - 1, hi + 1, 20000), edges, np.nextafter(edges, -np.inf),
                            np.nextafter(edges, np.inf), [np.nan, np.inf, -np.inf]])
        # histogram2d semantics
        want = np.searchsorted(edges, v, side="right") - 1
        want[v == edges[-1]] = nb - 1
        want[(want < 0) | (want > nb - 1) | np.isnan(v)] = -1
        assert np.array_equal(H.bins(v, lo, hi, nb, True), want)
        # digitize - 1 semantics
        want = np.digitize(v, edges) - 1
        want[(want < 0) | (want > nb - 1)] = -1
        assert np.array_equal(H.bins(v, lo, hi, nb, False), want)


def test_histogram_via_bins_equals_reference(H, golden):
    g = golden("g4_optics")
    for tag in ("shadow_single", "schlieren_DF", "refracto_incoherent"):
        r = H.optics(g["r0"], O.chain(tag))
        for bs in (25, 8):
            nx, ny = 3448 // bs, 2574 // bs
            ix, iy = H.bins(r[0], -9.0, 9.0, nx, True), H.bins(r[2], -6.75, 6.75, ny, True)
            ok = (ix >= 0) & (iy >= 0)
            Hh = np.zeros((ny, nx))
            np.add.at(Hh, (iy[ok], ix[ok]), 1.0)
            assert np.array_equal(Hh, g[f"{tag}_H{bs}"]), (tag, bs)


def test_philox_known_answer_and_beam_statistics(H):
    # Random123 known-answer vectors for philox4x32-10
    assert list(H.philox((0, 0), (0, 0, 0, 0))) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(H.philox((0xffffffff, 0xffffffff), (0xffffffff,) * 4)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(H.philox((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344))) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    n, R, div, ext = 200000, 4e-3, 5e-5, 10e-3
    for bt in (0, 1):                                    # fold(U+U) and power(2) radial laws
        s0 = H.beam(bt, 2, R, R, div, -ext, 7, 0, n)
        r = np.hypot(s0[0], s0[1]) / R
        assert r.max() <= 1.0 and np.all(s0[2] == -ext)
        # both laws have pdf 2r on [0,1] (uniform over the disc): mean 2/3, E[r^2] = 1/2
        assert abs(r.mean() - 2 / 3) < 3e-3 and abs((r ** 2).mean() - 0.5) < 3e-3
        ang = np.arctan2(s0[1], s0[0])
        assert abs(np.cos(ang).mean()) < 5e-3 and abs(np.sin(2 * ang).mean()) < 5e-3
        v = s0[3:6]
        assert np.allclose(np.linalg.norm(v, axis=0), C_LIGHT, rtol=1e-12)
        chi_abs = np.arcsin(np.hypot(v[0], v[1]) / C_LIGHT)          # |chi|, chi ~ divergence * N(0,1)
        assert abs(np.sqrt((chi_abs ** 2).mean()) / div - 1.0) < 0.01
    # partition invariance: ray i depends on (seed, i) only
    a = H.beam(1, 2, R, R, div, -ext, 7, 1000, 50)
    b = H.beam(1, 2, R, R, div, -ext, 7, 0, 1050)[:, 1000:]
    assert np.array_equal(a, b)


def test_rk4_reflecting_rays(H):
    """Over-critical density ramp: oblique rays turn around inside the plasma and leave through the face they entered
    (v_z changes sign, cells are crossed downwards, early exit must not fire while a ray outside is heading back in).
    Product RK4 against the loop around the reference RHS, with and without early exit."""
    n = 40
    x = np.linspace(-1e-3, 1e-3, n); z = np.linspace(-2e-3, 2e-3, 2 * n)
    omega = omega_of(1064e-9)
    nc = 3.14207787e-4 * omega ** 2
    _, _, ZZ = np.meshgrid(x, x, z, indexing="ij")
    ne = 1.6 * nc * np.clip((ZZ + 2e-3) / 4e-3, 0, 1)                       # 0 -> 1.6 n_c along z
    d = O.Domain(x, x, z, 2e-3)
    d.external_ne(ne)
    d.calc_dndr(1064e-9)
    f = H.field(ne, x, x, z, omega)
    rng = np.random.default_rng(4)
    N = 300
    s0 = np.zeros((9, N))
    s0[0], s0[1] = rng.uniform(-3e-4, 3e-4, N), rng.uniform(-3e-4, 3e-4, N)
    s0[2] = -2e-3 - 1e-5 * rng.random(N)                                     # start just outside, heading in
    th = rng.uniform(0.02, 0.1, N); ph = rng.uniform(0, 2 * np.pi, N)
    s0[3], s0[4], s0[5] = C_LIGHT * np.sin(th) * np.cos(ph), C_LIGHT * np.sin(th) * np.sin(ph), C_LIGHT * np.cos(th)
    s0[6] = 1.0
    h = 0.5 * (z[1] - z[0]) / C_LIGHT
    steps = 700
    for early in (False, True):
        ref, ref_steps = d.solve_rk4(s0, steps, h=h, early_exit=early)
        got, got_steps = f.rk4(s0, steps, h, early=early)
        assert np.mean(ref[5] < 0) > 0.9                                      # the rays did turn around
        assert rel_err(got[:6], ref[:6]) < 1e-9
        if early:
            assert np.array_equal(got_steps, ref_steps) and ref_steps.max() < steps


def test_fresnel_step_matches_reference(H, golden):
    """csrc/fresnel_core.h (host build) against g7 = the reference's fresnel_integral.py run unmodified."""
    from harness import Fresnel
    from scipy.signal.windows import tukey
    from scipy.spatial import Delaunay
    F, g = Fresnel(H), golden("g7_fresnel")
    for M in (1, 2, 5, 96, 360, 481):
        for alpha in (0.0, 0.4, 0.77, 1.0):
            assert np.allclose(F.window(M, alpha), tukey(M, alpha), rtol=0, atol=2e-16 if alpha < 1 else 2e-15), (M, alpha)
    a = np.arange(9)
    for n_pad in (3, 8, 20):
        assert np.array_equal(a[F.reflect(9, -n_pad, 9 + n_pad)], np.pad(a, n_pad, mode="reflect"))
    assert np.array_equal(F.reflect(1, -3, 4), np.zeros(7))
    # scattered rays -> grids on the Delaunay triangulation (Qhull on the host, as inside LinearNDInterpolator)
    r0, x, y = g["r0"], g["x"], g["y"]
    tri = Delaunay(np.stack([r0[0], r0[2]], axis=1)).simplices
    ph, am = F.scatter_to_grid(r0[0], r0[2], [g["phase"], g["amp"]], tri, x, y)
    assert np.array_equal(ph == 0.0, g["phase_grid"] == 0.0)                 # same nodes fall outside the hull
    assert np.abs(ph - g["phase_grid"]).max() <= 1e-11 * np.abs(g["phase_grid"]).max()
    assert np.abs(am - g["amp_grid"]).max() <= 1e-11
    # pad + window (fused with U0 = amp exp(-i phase)), transfer function, crop
    U0 = g["amp_grid"] * np.exp(-1j * g["phase_grid"])
    Lxy = (float(g["Lx"]), float(g["Ly"]))
    for pf in (2, 1):
        prep = F.prepare(g["amp_grid"], g["phase_grid"], pad=pf)
        assert np.abs(prep[::7, ::5] - g["prep_pf%d_sub" % pf]).max() <= 1e-15
        assert np.abs(F.prepare(U0, pad=pf) - prep).max() <= 1e-15          # complex-input mode
        out = F.propagate(prep, Lxy, float(g["lwl"]), float(g["z"]), U0.shape, pad=pf)
        ref = g["out_pf%d" % pf]
        assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    sigma = 150e-6 / (2 * np.sqrt(2 * np.log(2)))
    out = F.propagate(F.prepare(U0), Lxy, float(g["lwl"]), float(g["z"]), U0.shape, sigma=sigma)
    assert np.abs(out - g["out_lanex"]).max() <= 1e-12 * np.abs(g["out_lanex"]).max()
    # end to end from the device-side interpolation
    out = F.propagate(F.prepare(am, ph), Lxy, float(g["lwl"]), float(g["z"]), U0.shape)
    assert np.abs(out - g["out_pf2"]).max() <= 1e-9 * np.abs(g["out_pf2"]).max()


def test_scatter_to_grid_edge_cases(H):
    """Nodes on vertices and edges, a non-uniform grid, nodes outside the hull, duplicate-free degenerate input."""
    from harness import Fresnel
    from scipy.interpolate import LinearNDInterpolator
    from scipy.spatial import Delaunay
    F = Fresnel(H)
    rng = np.random.default_rng(5)
    gx = np.sort(np.concatenate([np.linspace(-1, 1, 21), rng.uniform(-1, 1, 12)]))
    gy = np.linspace(-0.8, 0.8, 17)
    # sample positions include exact grid nodes (vertex hits) and points on grid lines (edge hits)
    XX, YY = np.meshgrid(gx[::4], gy[::3])
    pts = np.concatenate([np.stack([XX.ravel(), YY.ravel()], 1), rng.uniform(-0.9, 0.9, (400, 2))])
    v = np.sin(3 * pts[:, 0]) * np.cos(2 * pts[:, 1]) + pts[:, 0]
    tri = Delaunay(pts).simplices
    out = F.scatter_to_grid(pts[:, 0], pts[:, 1], [v], tri, gx, gy, fill=-7.0)[0]
    ref = LinearNDInterpolator(pts, v, fill_value=-7.0)(*np.meshgrid(gx, gy))
    assert np.array_equal(out == -7.0, ref == -7.0)
    assert np.abs(out - ref).max() <= 1e-12
    # a linear function is reproduced exactly (to rounding) inside the hull whatever the triangulation
    lin = 2.0 * pts[:, 0] - 3.0 * pts[:, 1] + 0.5
    out = F.scatter_to_grid(pts[:, 0], pts[:, 1], [lin], tri, gx, gy, fill=np.nan)[0]
    GX, GY = np.meshgrid(gx, gy)
    m = ~np.isnan(out)
    assert m.sum() > 0.7 * m.size and np.abs(out[m] - (2.0 * GX - 3.0 * GY + 0.5)[m]).max() <= 1e-13
    # no triangles at all: everything is fill
    out = F.scatter_to_grid(pts[:3, 0], pts[:3, 1], [v[:3]], np.zeros((0, 3), np.int32), gx, gy, fill=1.5)[0]
    assert np.all(out == 1.5)


def test_fresnel_primitives_property(H):
    """reflect_idx == np.pad('reflect') and tukey_w == scipy's window for arbitrary sizes (hypothesis)."""
    from harness import Fresnel
    from hypothesis import given, settings, strategies as st
    from scipy.signal.windows import tukey
    F = Fresnel(H)

    @settings(max_examples=150, deadline=None)
    @given(n=st.integers(1, 40), pad=st.integers(0, 130))
    def reflect(n, pad):
        a = np.arange(n)
        assert np.array_equal(a[F.reflect(n, -pad, n + pad)], np.pad(a, pad, mode="reflect"))

    @settings(max_examples=150, deadline=None)
    @given(M=st.integers(1, 700), alpha=st.one_of(st.just(0.0), st.floats(1e-3, 1.0)))      # scipy itself overflows for denormal alpha
    def window(M, alpha):
        assert np.allclose(F.window(M, alpha), tukey(M, alpha), rtol=0, atol=4e-15)
    reflect()
    window()


def test_tsit5_pid_flavour(H, golden):
    """method='tsit5' (the current generation's diffrax solve, propagator.py:533-599; PARITY UNPINNED -- no jax here):
    the product arithmetic (Nystrom-free 7-row form on the cache-free RHS) against the oracle's plain 9-vector
    restatement of the same published method: identical accept / reject sequences, states to 1e-9 -- at upstream's
    shipped controller setting (rtol 1, atol 1e-5: the box in ~a dozen steps) and at its 'intended' 1e-7 / 1e-9."""
    g = golden("g2_expcos")
    ext = float(g["extent"])
    f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(float(g["lwl"])), phase=True, f64=True)
    o = O.Domain(g["x"], g["y"], g["z"], ext, phaseshift=True)
    o.external_ne(g["ne"]); o.calc_dndr(float(g["lwl"]))
    T = np.sqrt(8.0) * ext / C_LIGHT
    s0 = g["s0"][:, :24]
    for rtol, atol, tol in ((1.0, 1e-5, 1e-9), (1e-3, 1e-6, 1e-7)):
        sf, att, acc = f.tsit5(s0, T, T / 2, rtol, atol)
        sf_o, att_o, acc_o = O.solve_tsit5_per_ray(o, s0, rtol, atol)
        assert np.array_equal(att, att_o) and np.array_equal(acc, acc_o)              # same accept / reject sequence
        assert np.max(np.abs(sf[:3] - sf_o[:3])) < tol * ext and np.max(np.abs(sf[3:6] - sf_o[3:6])) < tol * C_LIGHT
        assert np.max(np.abs(sf[7] - sf_o[7])) < 10 * tol * max(1.0, np.abs(sf_o[7]).max())
    assert att.min() >= 12                                           # shipped setting: a dozen steps through the whole box
    # 1e-7 / 1e-9: as for RK45 (above) the error norm in the free-flight legs is rounding noise, sequences fork between any
    # two implementations; agreement at the level of the tolerance and a comparable amount of work
    sf, att, acc = f.tsit5(s0, T, T / 2, 1e-7, 1e-9)
    sf_o, att_o, acc_o = O.solve_tsit5_per_ray(o, s0, 1e-7, 1e-9)
    assert abs(int(att.sum()) - int(att_o.sum())) < 0.15 * att_o.sum() and att.mean() > 40
    assert np.max(np.abs(sf[:3] - sf_o[:3])) < 2e-6 * ext and np.max(np.abs(sf[3:6] - sf_o[3:6])) < 2e-6 * C_LIGHT
    # order conditions of the tableau (Tsitouras 2011): row sums, quadrature and tree conditions up to order 5
    c, A, b, bt = O.TSIT5_C, O.TSIT5_A, O.TSIT5_B, O.TSIT5_BT
    assert np.abs(A.sum(1) - c).max() < 1e-15
    for k in range(5):
        assert abs(b @ c ** k - 1 / (k + 1)) < 1e-15
    assert abs(b @ A @ c - 1 / 6) < 1e-15 and abs(b @ A @ A @ c - 1 / 24) < 1e-15 and abs(b @ A @ A @ A @ c - 1 / 120) < 1e-15
    assert abs((b * c) @ A @ c - 1 / 8) < 1e-15 and abs(b @ A @ c ** 3 - 1 / 20) < 1e-15 and abs(b @ ((A @ c) ** 2) - 1 / 20) < 1e-15
    assert abs(bt.sum()) < 1e-15 and all(abs(bt @ c ** k) < 1e-15 for k in (1, 2, 3)) and abs(bt @ c ** 4) > 1e-4


def test_rhs_direct_matches_reference_and_cached_path(H, golden):
    """The cache-free evaluation the adaptive integrators use (rhs_direct: guessed cell, table entries and corners in one
    round trip, verified, reference's sum over corners of value x weight product): (a) against the reference's own dsdt
    (golden g1: identical in/out-of-bounds pattern incl. points exactly on the rounded end nodes, <= 1e-11; phase lane to the
    oracle's own rounding floor); (b) against the cell-cached evaluation of the fixed-step kernel to rounding; (c) on
    non-uniform axes, on nodes, one ulp either side of nodes, outside the grid, and NaN in -> NaN out."""
    g = golden("g1_rhs")
    for ph in (False, True):
        f = H.field(g["ne"], g["x"], g["y"], g["z"], omega_of(float(g["lwl"])), phase=ph, f64=ph)
        out, ref = f.rhs_direct(g["s"]), g["dsdt_phase%d" % ph]
        assert np.array_equal(out[:3], ref[:3])
        assert np.array_equal(out[3:6] == 0, ref[3:6] == 0)
        assert rel_err(out[3:6], ref[3:6], floor=1e3) < 1e-11
        cached = f.rhs(g["s"])
        assert np.max(np.abs(out[3:6] - cached[3:6])) < 1e-13 * np.abs(cached[3:6]).max()
        if ph:
            assert np.all(np.abs(out[7] - ref[7]) <= 1e-15 * f.omega + 1e-12 * np.abs(ref[7]))
    rng = np.random.default_rng(8)
    x = np.cumsum(rng.uniform(0.5, 2.0, 19)); x = (x - x.mean()) * 1e-3 / 3
    y = np.sort(rng.uniform(-4e-3, 4e-3, 15)); z = np.linspace(-1, 1, 23) ** 3 * 8e-3
    ne = 1e25 * (1 + 0.5 * rng.random((19, 15, 23)))
    d = O.Domain(x, y, z, 8e-3)
    d.external_ne(ne)
    d.calc_dndr(1064e-9)
    f = H.field(ne, x, y, z, omega_of(1064e-9))
    ax = [np.float64(np.float32(a)) for a in (x, y, z)]
    n = 6000
    s = np.zeros((9, n))
    s[0], s[1], s[2] = rng.uniform(x[0] * 1.1, x[-1] * 1.1, n), rng.uniform(-4.4e-3, 4.4e-3, n), rng.uniform(-8.5e-3, 8.5e-3, n)
    for k in range(3):                                             # nodes, and one ulp either side of them
        m = len(ax[k])
        s[k, 100 * k:100 * k + m] = ax[k]
        s[k, 1000 + 100 * k:1000 + 100 * k + m] = np.nextafter(ax[k], np.inf)
        s[k, 2000 + 100 * k:2000 + 100 * k + m] = np.nextafter(ax[k], -np.inf)
    s[3:6] = 1e8
    ref = d.dsdt(0.0, s.ravel().copy()).reshape(9, -1)
    out = f.rhs_direct(s)
    assert np.array_equal(out[3:6] == 0, ref[3:6] == 0)
    assert np.max(np.abs(out[3:6] - ref[3:6])) < 1e-11 * np.abs(ref[3:6]).max()
    s[0, 5] = np.nan
    assert np.all(np.isnan(f.rhs_direct(s)[3:6, 5]))
