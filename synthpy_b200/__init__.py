"""synthpy_b200 -- B200-native implementation of synthPy's ray-propagation hot path.

Same call shape as the reference's ``src/simulator`` package::

    from synthpy_b200 import domain as d, beam as b, propagator as p, diagnostics as diag
    dom = d.ScalarDomain(lengths, dims, ne_type="test_exponential_cos")
    rays = b.Beam(Np, beam_size, divergence, ne_extent)
    rf, Jf, duration = p.solve(rays.s0, dom, probing_extent)
    sh = diag.Shadowgraphy(lwl, rf); sh.single_lens_solve(); sh.histogram(bin_scale=1); sh.H

The compute is hand-written sm_100a CUDA behind a C ABI (include/synthpy_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError if the CUDA library is not built)
from . import beam, diagnostics, domain, engine, out_of_core, propagator  # noqa: F401

__all__ = ["beam", "diagnostics", "domain", "engine", "out_of_core", "propagator"]
