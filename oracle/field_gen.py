"""TEST INFRASTRUCTURE (checker / CPU-baseline input only): CPU restatement of the reference's turbulent-field
generator, so that bench.py's ``--impl reference`` arm never imports the product package.

Follows ``gaussian3D.domain_fft`` (reference: src/field_generator/gaussian3D.py:215-271) and the
``ne = ne0 + dne * f`` recipe of examples/.../turb_gen.py:36-50.  The complex noise is drawn either with NumPy's
legacy global RNG in the reference's own order (noise='numpy': the reference's realisation for a given
``np.random.seed``) or with torch's CPU generator (noise='torch': the realisation bench.py's GPU arm uses, whose
noise is also drawn on the CPU), and the inverse FFT runs on the host.
"""
import numpy as np


def domain_fft(k_func, l_max, l_min, extent, res, factor=1, *, noise="numpy", seed=0):
    """(2res, 2res, int(2res*factor)) float64 field normalised to max |f| = 1."""
    nx = ny = 2 * res
    nz = int(2 * res * factor)
    dx = extent / res
    kx = ky = 2 * np.pi * np.fft.fftfreq(nx, d=dx)
    kz = 2 * np.pi * np.fft.fftfreq(nz, d=dx)
    kxx, kyy, kzz = np.meshgrid(kx, ky, kz, copy=False)          # 'xy' indexing, gaussian3D.py:238
    k = np.sqrt(kxx ** 2 + kyy ** 2 + kzz ** 2, dtype=np.float32)
    del kxx, kyy, kzz
    k_min, k_max = 2 * np.pi / l_max, 2 * np.pi / l_min
    S = np.zeros_like(k)
    mask = (k >= k_min) & (k <= k_max)
    S[mask] = k_func(k[mask])
    del mask
    shape = k.shape
    del k
    if noise == "numpy":
        w = np.random.normal(0, 1, shape) + 1j * np.random.normal(0, 1, shape)
    else:
        import torch
        gen = torch.Generator(device="cpu").manual_seed(int(seed))
        re = torch.randn(shape, generator=gen, dtype=torch.float64).numpy()
        im = torch.randn(shape, generator=gen, dtype=torch.float64).numpy()
        w = re + 1j * im
        del re, im
    w *= np.sqrt(S)
    del S
    import scipy.fft as sfft
    field = sfft.ifftn(w, workers=-1, overwrite_x=True).real
    return field / np.abs(field).max()


def kolmogorov(k):
    return k ** (-11.0 / 3.0)


def turbulent_ne(res, *, ne0=1e25, dne=9e24, l_max=1, l_min=0.01, extent=5, noise="torch", seed=1):
    """ne = ne0 + dne f on a (2 res)^3 grid: the C2 / C5 field of BASELINE.json."""
    return ne0 + dne * domain_fft(kolmogorov, l_max, l_min, extent, res, 1, noise=noise, seed=seed)
