"""TEST INFRASTRUCTURE (checker only): fan the oracle's integrators out over the host cores.

The oracle (oracle/synthpy_oracle.py) is single-threaded NumPy/SciPy; on the benchmarked 512^3 field a
20 000-ray fixed-step solve takes ~2 minutes on one core.  Rays are independent (every reference driver chunks
them: examples/jobs/run_scripts/pvti_trace_multiprocess.py:102-134), so chunks of the same ``s0`` go to a fork
pool that shares the prepared ``Domain`` copy-on-write.  Results are concatenated in ray order: identical to the
single-process call for the fixed-step and per-ray solvers (neither couples rays).

Only tests/, __graft_entry__.smoke() and bench.py's checker / CPU legs may import this module.
"""
import multiprocessing as mp
import os

import numpy as np

_STATE = {}


def _init():
    try:                                     # one BLAS thread per worker: the pool already uses every core
        from threadpoolctl import threadpool_limits
        _STATE["tp"] = threadpool_limits(1)
    except Exception:
        pass


def _rk4(args):
    lo, hi, n_steps, h, early = args
    return _STATE["dom"].solve_rk4(_STATE["s0"][:, lo:hi], n_steps, h=h, early_exit=early)


def _per_ray(args):
    lo, hi, rtol, atol = args
    return _STATE["dom"].solve_per_ray(_STATE["s0"][:, lo:hi], rtol=rtol, atol=atol)


def _chunks(n, workers, min_chunk):
    per = max(min_chunk, -(-n // max(1, workers)))
    return [(a, min(n, a + per)) for a in range(0, n, per)]


def _run(fn, dom, s0, jobs, workers):
    _STATE["dom"], _STATE["s0"] = dom, np.ascontiguousarray(s0, dtype=np.float64)
    try:
        if workers <= 1 or len(jobs) == 1:
            return [fn(j) for j in jobs]
        with mp.get_context("fork").Pool(min(workers, len(jobs)), initializer=_init) as pool:
            return pool.map(fn, jobs)
    finally:
        _STATE.pop("dom", None); _STATE.pop("s0", None)


def solve_rk4(dom, s0, n_steps, h=None, early_exit=False, workers=None):
    """``Domain.solve_rk4`` over a pool: returns (9xN state, steps per ray)."""
    workers = workers or os.cpu_count() or 1
    jobs = [(a, b, n_steps, h, early_exit) for a, b in _chunks(s0.shape[1], workers, 64)]
    res = _run(_rk4, dom, s0, jobs, workers)
    return np.concatenate([r[0] for r in res], axis=1), np.concatenate([r[1] for r in res])


def solve_per_ray(dom, s0, rtol=1e-3, atol=1e-6, workers=None):
    """``Domain.solve_per_ray`` over a pool: returns (9xN state, nfev per ray)."""
    workers = workers or os.cpu_count() or 1
    jobs = [(a, b, rtol, atol) for a, b in _chunks(s0.shape[1], workers, 4)]
    res = _run(_per_ray, dom, s0, jobs, workers)
    return np.concatenate([r[0] for r in res], axis=1), np.concatenate([r[1] for r in res])
