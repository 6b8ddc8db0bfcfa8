"""Band-limited Gaussian random field on the GPU: restates ``gaussian3D.domain_fft``
(reference: src/field_generator/gaussian3D.py:215-271) so that 512^3 / 1024^3 turbulent n_e grids can be built
where they are used.  Input producer for the ray path, not part of it; torch.fft is the
library FFT here (plumbing), the ray kernels are in csrc/.

noise='numpy' draws the complex noise with NumPy's legacy global RNG in the reference's order (same field as
the reference for the same ``np.random.seed``); noise='torch' draws it with torch's (multi-threaded) CPU generator -- fast enough for 1024^3.
"""
import numpy as np
import torch


def domain_fft(k_func, l_max, l_min, extent, res, factor=1, *, noise="numpy", seed=0, device="cuda"):
    """Returns the (2res, 2res, int(2res*factor)) field as a float64 torch tensor on ``device``, normalised to
    max |f| = 1.  k_func maps a float32 torch tensor of |k| to the power spectrum."""
    nx = ny = 2 * res
    nz = int(2 * res * factor)
    dx = extent / res
    kx = 2 * np.pi * np.fft.fftfreq(nx, d=dx)
    kz = 2 * np.pi * np.fft.fftfreq(nz, d=dx)
    dev = torch.device(device)
    kxt, kzt = torch.from_numpy(kx).to(dev), torch.from_numpy(kz).to(dev)
    # np.meshgrid(kx, ky, kz) uses 'xy' indexing: axis 0 <- ky, axis 1 <- kx   (gaussian3D.py:238)
    k2 = kxt[None, :, None] ** 2 + kxt[:, None, None] ** 2 + kzt[None, None, :] ** 2
    k = torch.sqrt(k2.to(torch.float32))                       # np.sqrt(..., dtype=np.float32)
    del k2
    k_min, k_max = 2 * np.pi / l_max, 2 * np.pi / l_min
    mask = (k >= k_min) & (k <= k_max)
    S = torch.zeros_like(k)
    S[mask] = k_func(k[mask])
    del mask, k
    shape = (ny, nx, nz)
    if noise == "numpy":
        re = torch.from_numpy(np.random.normal(0, 1, shape)).to(dev)
        im = torch.from_numpy(np.random.normal(0, 1, shape)).to(dev)
    else:
        # drawn with torch's CPU generator so that the realisation does not depend on where the FFT runs
        gen = torch.Generator(device="cpu").manual_seed(int(seed))
        re = torch.randn(shape, generator=gen, dtype=torch.float64).to(dev)
        im = torch.randn(shape, generator=gen, dtype=torch.float64).to(dev)
    spec = torch.complex(re, im) * torch.sqrt(S).to(torch.float64)
    del re, im, S
    field = torch.fft.ifftn(spec).real
    del spec
    return field / field.abs().max()


def kolmogorov(k):
    return k ** (-11.0 / 3.0)                                   # examples/.../turb_gen.py:36-50


def turbulent_ne(res, *, ne0=1e25, dne=9e24, l_max=1, l_min=0.01, extent=5, noise="torch", seed=1, device="cuda"):
    """ne = ne0 + dne * f with f the k^-11/3 field of ``domain_fft`` on a (2res)^3 grid (BASELINE config C2/C5)."""
    f = domain_fft(kolmogorov, l_max, l_min, extent, res, 1, noise=noise, seed=seed, device=device)
    return ne0 + dne * f
