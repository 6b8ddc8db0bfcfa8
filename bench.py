#!/usr/bin/env python
"""Benchmark of the ray-propagation hot path (BASELINE.json metric: rays.steps/s, device-timed, max over ranks).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA arm
  python bench.py --impl reference ...                           the reference's CPU algorithm (oracle port) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...              one rank per GPU, rays sharded, field replicated,
                                                                 detector images combined by one NCCL all-reduce

Workload (BASELINE.json configs[1], "C2"): 1e7 rays per GPU through a 512^3 turbulent (k^-11/3 power spectrum,
field_generator.domain_fft) n_e field, lambda = 1064 nm, box 10 x 10 x 20 mm, circular beam r = 5 mm, divergence
5e-5; fixed-step RK4 with two steps per cell and early exit; shadowgraphy (two-lens) + dark-field schlieren
images at full 3448 x 2574 resolution fused into the propagation kernel.  One "step" = one pass of that whole
bundle.  Rays are generated on the device (Philox), so nothing but the replicated field is resident input.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LWL = 1064e-9
LENGTHS = (10e-3, 10e-3, 20e-3)
EXTENT = 10e-3
BEAM_R, BEAM_DIV = 5e-3, 5e-5
BYTES_PER_RAY_STEP = 512          # 4 RHS evaluations x 8 corners x 16 B
C_LIGHT = 299792458.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--rays", type=float, default=1e7, help="rays per GPU per step")
    ap.add_argument("--bin-scale", type=int, default=1)
    ap.add_argument("--no-sort", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-rays-per-worker", type=int, default=4000)
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--ds-frac", type=float, default=0.5, help="RK4 step as a fraction of the cell size along the probing axis")
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4", "C5"],
                    help="C2 shadowgraphy+schlieren (default, the headline); C3 interferometry with phase accumulation; "
                         "C4 refractometry + knife-edge schlieren with adaptive RK45; C5 shadowgraphy on a 1024^3 field with "
                         "1e9 rays over 8 GPUs (--grid 1024 --rays 1.25e8 unless given)")
    ap.add_argument("--bundle", action="store_true", help="C4: one step size per 32-ray bundle instead of per ray")
    ap.add_argument("--rtol", type=float, default=1e-3)
    ap.add_argument("--atol", type=float, default=1e-6)
    a = ap.parse_args()
    if a.workload == "C5":                      # BASELINE configs[4]: 1e9 rays through 1024^3, sharded over the GPUs
        if "--grid" not in sys.argv:
            a.grid = 1024
        if "--rays" not in sys.argv:
            a.rays = 1.25e8
        a.no_cpu_baseline = True               # the CPU port on a 1024^3 grid needs tens of GB and minutes: C2 carries it
    return a


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_ne(grid, device):
    from synthpy_b200 import field_generator as fg
    return fg.turbulent_ne(grid // 2, noise="torch", seed=1, device=device)      # ne = 1e25 + 9e24 f, (grid)^3


# ------------------------------------------------------------------------------------------------ CPU arms
_CPU = {}


def _cpu_worker(args):
    """One chunk of rays through the oracle's restatement of the shipped solver (joint RK45, full_solver.py:391)."""
    seed, n = args
    from oracle import synthpy_oracle as O
    dom = _CPU["dom"]
    rng = np.random.RandomState(seed)
    s0 = O.init_beam(n, BEAM_R, BEAM_DIV, EXTENT, "circular", "z", rng=rng)
    t0 = time.perf_counter()
    sf, sol = dom.solve_joint(s0, return_stats=True)
    rf, _ = O.ray_to_jones(sf, EXTENT)
    r = O.run_chain(rf, O.chain("shadow_two"))
    O.histogram(r, bin_scale=1)
    return n * (sol.nfev - 2) / 6.0, time.perf_counter() - t0


def _cpu_worker_rk4(args):
    """The GPU arm's own integrator on the CPU: fixed-step RK4 (same step, early exit) around the reference's RHS."""
    seed, n, h, n_steps = args
    from oracle import synthpy_oracle as O
    dom = _CPU["dom"]
    s0 = O.init_beam(n, BEAM_R, BEAM_DIV, EXTENT, "circular", "z", rng=np.random.RandomState(seed))
    t0 = time.perf_counter()
    sf, steps = dom.solve_rk4(s0, n_steps, h=h, early_exit=True)
    rf, _ = O.ray_to_jones(sf, EXTENT)
    O.histogram(O.run_chain(rf, O.chain("shadow_two")), bin_scale=1)
    return float(steps.sum()), time.perf_counter() - t0


def _pool_init():
    """One BLAS/OpenMP thread per worker process: the pool already uses every core (the reference's own
    config pins threads to 1 as well, src/simulator/config.py:80-122); without this np.dot inside solve_ivp
    oversubscribes the box and timings become erratic."""
    try:
        from threadpoolctl import threadpool_limits
        _CPU["tp"] = threadpool_limits(1)
    except Exception:
        pass


def cpu_setup(ne_host, grid):
    from oracle import synthpy_oracle as O
    x, y, z = (np.linspace(-L / 2, L / 2, grid) for L in LENGTHS)
    dom = O.Domain(x, y, z, EXTENT)
    dom.external_ne(ne_host)
    dom.calc_dndr(LWL)
    dom.ne = None
    _CPU["dom"] = dom


def cpu_pass(pool, cores, rays_per_worker, seed0):
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, [(seed0 + i, rays_per_worker) for i in range(cores)])
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


def run_reference(a):
    """--impl reference: the reference's own CPU algorithm (NumPy/SciPy restatement in oracle/, pinned to the
    real reference by tests/golden) on all host cores, mirroring the reference's multiprocessing driver
    (examples/jobs/run_scripts/pvti_trace_multiprocess.py:102-134)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import torch
    cores = os.cpu_count() or 1
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    ne = build_ne(a.grid, dev).cpu().numpy()
    cpu_setup(ne, a.grid)
    del ne
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_pool_init) as pool:
        for w in range(a.warmup):
            cpu_pass(pool, cores, max(8, a.cpu_rays_per_worker // 8), 1000 + 100 * w)
        t0 = time.perf_counter()
        units = 0.0
        for k in range(a.steps):
            u, _ = cpu_pass(pool, cores, a.cpu_rays_per_worker, 5000 + 100 * k)
            units += u
        wall = time.perf_counter() - t0
    value = units / wall
    sample = (f"{cores} workers x {a.cpu_rays_per_worker} rays per step, joint RK45 (rtol 1e-3, atol 1e-6) + two-lens "
              f"shadowgraphy + histogram, same {a.grid}^3 turbulent field")
    line = {"impl": "reference", "metric": "rays*steps/s", "value": value, "unit": "rays*steps/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * wall / max(1, a.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a),
            "cpu_baseline": {"value": value, "unit": "rays*steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rays*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(a):
    diag = {"C2": "shadowgraphy(two-lens) + schlieren(DF)", "C3": "interferometry(two-lens, phase accumulation, reference beam)",
            "C4": "refractometry(incoherent) + knife-edge schlieren", "C5": "shadowgraphy(two-lens)"}[a.workload]
    integ = (f"rk4, ds = {a.ds_frac:g} cell, early exit" if a.workload != "C4" else
             f"rk45 {'per 32-ray bundle' if a.bundle else 'per ray'} (SciPy controller), rtol {a.rtol:g} atol {a.atol:g}, early exit")
    return {"workload": f"{a.workload}: {int(a.rays):d} rays/GPU through a {a.grid}^3 turbulent (k^-11/3) n_e field, "
                        f"{diag} at bin_scale {a.bin_scale}",
            "grid": a.grid, "rays_per_gpu": int(a.rays), "integrator": integ,
            "precision": "fp32" if a.fp32 else "fp64", "field_bytes": 16 * a.grid ** 3,
            "l2_policy": f"inputs larger than L2 (packed field {16 * a.grid ** 3 / 1e9:.2f} GB vs 126 MB L2); no flush needed",
            "rays": "generated on device (Philox4x32-10), sorted into cell-column bundles" if not a.no_sort else
                    "generated on device, unsorted"}


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from synthpy_b200 import _lib, beam as B, diagnostics as D, domain as Dm, engine, propagator as P

    n_rays = int(a.rays)
    ne = build_ne(a.grid, "cuda")
    dom = Dm.ScalarDomain(LENGTHS, a.grid)
    dom.external_ne(ne)
    dom.device_field(LWL)
    del ne
    torch.cuda.empty_cache()
    kw = dict(lwl=LWL, method="rk4", precision="fp32" if a.fp32 else "fp64", sort=not a.no_sort,
              ds=a.ds_frac * dom.cell_size())
    if a.workload == "C2":
        specs = [D.spec("shadow_two", bin_scale=a.bin_scale), D.spec("schlieren_DF", bin_scale=a.bin_scale, R_stop=1)]
    elif a.workload == "C5":       # BASELINE configs[4]: shadowgraphy only, 1024^3 field replicated on every GPU
        specs = [D.spec("shadow_two", bin_scale=a.bin_scale)]
    elif a.workload == "C3":       # BASELINE configs[2]: phase accumulation + 2-D interferogram (reference beam 10 fringes, 20 deg)
        specs = [D.spec("interf_two", bin_scale=a.bin_scale, interferogram=True, wavelength=LWL, ref_beam=(10, 20))]
    else:                          # BASELINE configs[3]: refractometry + knife-edge schlieren, adaptive RK45 (SciPy controller)
        specs = [D.spec("refracto_incoherent", bin_scale=a.bin_scale),
                 D.spec("schlieren_knife", bin_scale=a.bin_scale, offset=0.1, axis=2, direction=1)]
        kw.update(method="rk45_bundle" if a.bundle else "rk45", rtol=a.rtol, atol=a.atol, max_steps=1000000)
        kw.pop("ds")
    beam = B.Beam(n_rays, BEAM_R, BEAM_DIV, EXTENT, device=True, seed=2, beam_type="circular")

    def one_pass(rays, sync=False):
        for s in specs:
            s.image.zero_()
        st, _ = P.solve_and_image(dom, rays, EXTENT, specs, n_rays=n_rays, ray_offset=rank * n_rays, sync=False, **kw)
        if world > 1:                         # the path's one exchange: sum of detector images
            for s in specs:
                for t_ in s.image.tensors():
                    dist.all_reduce(t_, op=dist.ReduceOp.SUM)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, a.warmup)):
        st = one_pass(beam)
    barrier()
    steps_per_pass = engine.stats_dict(st)["ray_steps"]
    engine.propagate_kernel_ms()                       # reset the event log
    launches0 = _lib.launch_count()
    clocks = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        st = one_pass(beam)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if clocks else None
    launches = _lib.launch_count() - launches0
    kms, klaunch = engine.propagate_kernel_ms()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot_steps = torch.tensor([float(steps_per_pass) * a.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot_steps, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    value = float(tot_steps.item()) / (ms * 1e-3)

    # ---- e2e: host ray bundle in pinned memory -> H2D -> fused trace -> images D2H, every step
    e2e = None
    if not a.no_e2e:
        s0_host = beam.materialise(n_rays, rank * n_rays).cpu().pin_memory()
        outs = [torch.empty(s.image.tensors()[0].shape, dtype=s.image.tensors()[0].dtype).pin_memory() for s in specs]

        def e2e_pass(tok):
            st_ = one_pass(tok)
            for o, s in zip(outs, specs):
                o.copy_(s.image.tensors()[0], non_blocking=True)
            return st_
        # propagator.prefetch_rays: the host->device copy of step k+1's rays runs on a side stream while step k
        # propagates (two device buffers); every step's copy and image read-back is inside the timed region
        e2e_pass(P.prefetch_rays(s0_host))
        barrier()
        t0 = time.perf_counter()
        nxt = P.prefetch_rays(s0_host)
        for k in range(a.steps):
            cur, nxt = nxt, (P.prefetch_rays(s0_host) if k + 1 < a.steps else None)
            st_ = e2e_pass(cur)
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e_steps = engine.stats_dict(st_)["ray_steps"] * a.steps * world
        e2e = {"value": e_steps / float(tw.item()), "unit": "rays*steps/s",
               "h2d_bytes_per_step": int(s0_host.numel() * 8), "d2h_bytes_per_step": int(sum(o.numel() * 8 for o in outs)),
               "api": "propagator.prefetch_rays(s0_host_pinned) -> propagator.solve_and_image(domain, handle, ...) + image read-back"}
        del s0_host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    per_launch_steps = steps_per_pass * a.steps / max(1, klaunch)
    bytes_per_step = 768 if a.workload == "C4" else BYTES_PER_RAY_STEP      # DP5: 6 evaluations x 128 B; RK4: 4 x 128 B
    achieved = per_launch_steps * bytes_per_step / (kms / max(1, klaunch) * 1e-3) / 1e9 if kms > 0 else None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if achieved else None,
            "traffic": None, "kernel": "k_propagate<%s, %s>" % ("float" if a.fp32 else "double", "RK45" if a.workload == "C4" else "RK4"), "kernel_ms_per_launch": kms / max(1, klaunch),
            "kernel_share_of_step": kms / ms if ms > 0 else None, "peak_source": peak_src,
            "note": "achieved = algorithmic gather bytes (%d B per ray-step) / event-timed kernel duration; " % bytes_per_step +
                    "gathers are served mostly by L1/L2 (rays are bundled per cell column), so frac may exceed 1; "
                    "see profiles/ for dram__bytes and L2 hit rate"}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
    roof["bytes_per_ray_step"] = bytes_per_step
    roof["frac_of_nominal_8TBps"] = achieved / 8000.0 if achieved else None        # SURVEY.md 8d asks for both denominators
    if os.path.exists(traffic_file) and a.workload == "C2" and a.grid == 512 and n_rays == int(1e7) and not a.fp32:
        try:
            roof["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
        except Exception:
            pass
    prof = os.path.join(ROOT, "profiles", "r1_v34_k_propagate_ncu_full.json")
    if roof["traffic"] is not None and os.path.exists(prof):
        # what actually bounds the kernel: pipe utilisations from the committed ncu capture of this
        # same command -- static evidence, not measured in this run
        try:
            m = json.load(open(prof))
            pick = {"fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                    "l1tex_lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                    "l2_hit_pct": "lts__t_sector_hit_rate.pct", "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
                    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"}
            roof["ncu"] = {k: float(m[v][0]) for k, v in pick.items() if v in m}
            roof["ncu"]["source"] = "profiles/r1_v34_k_propagate_ncu_full.json"
        except Exception:
            pass
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a, dom)
    stats = engine.stats_dict(st)
    line = {"metric": "rays*steps/s", "value": value, "unit": "rays*steps/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(3, a.warmup), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if a.fp32 else "f64", "data": "synthetic", "config": workload_config(a),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
            "ray_steps_per_pass_per_gpu": int(steps_per_pass), "rays_binned_last_pass": stats["rays_binned"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(a, dom):
    """The oracle port of the shipped CPU solver, all host cores, bounded sample of the same workload."""
    import multiprocessing as mp
    import torch
    cores = os.cpu_count() or 1
    ne = dom.ne.cpu().numpy() if isinstance(dom.ne, torch.Tensor) else np.asarray(dom.ne)
    cpu_setup(ne, a.grid)
    del ne
    ctx = mp.get_context("fork")
    h = a.ds_frac * (LENGTHS[2] / (a.grid - 1)) / C_LIGHT
    n_steps = int(np.ceil(np.sqrt(8.0) * EXTENT / C_LIGHT / h))
    n4 = max(8, a.cpu_rays_per_worker // 16)
    with ctx.Pool(cores, initializer=_pool_init) as pool:
        cpu_pass(pool, cores, 8, 100)
        units, wall = cpu_pass(pool, cores, a.cpu_rays_per_worker, 200)
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker_rk4, [(300 + i, n4, h, n_steps) for i in range(cores)])
        wall4 = time.perf_counter() - t0
    return {"value": units / wall, "unit": "rays*steps/s", "cores": cores, "kind": "port",
            "sample": f"{cores} workers x {a.cpu_rays_per_worker} rays, joint RK45 (SciPy defaults) + two-lens "
                      f"shadowgraphy + histogram on the same {a.grid}^3 field; {wall:.1f} s wall",
            "same_integrator": {"value": sum(r[0] for r in res) / wall4, "unit": "rays*steps/s",
                                "sample": f"{cores} workers x {n4} rays, the GPU arm's fixed-step RK4 (ds = {a.ds_frac:g} cell, early "
                                          f"exit) around the reference's dsdt; {wall4:.1f} s wall"}}


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
