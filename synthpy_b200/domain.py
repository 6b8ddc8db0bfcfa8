"""Field container with the call shape of the reference's ``src/simulator/domain.py::ScalarDomain``.

Only what the ray path needs is kept: float32-rounded axes, the ``ne`` grid, analytic test profiles and
``external_*`` loaders (which, unlike upstream's frozen ``eqx.Module`` -- domain.py:310,453-461 -- work).
The memory-driven domain batching of the reference (domain.py:137-243, WIP upstream) is not needed: a
1024^3 packed field is 17 GB of a B200's 180 GB.  The packed device field is built lazily per wavelength.
"""
import numpy as np

from . import engine


class ScalarDomain:
    """Deviations from upstream's container that a caller should know:
      * ``region_count`` / ``auto_batching`` (domain.py:137-243, WIP upstream) are accepted and ignored -- one region, always
        (a warning is raised for region_count > 1); grids that do not fit in HBM are traced slab by slab with
        ``out_of_core.solve_out_of_core``, with results identical to one region;
      * the ``test_*`` profiles are evaluated on the float64 mesh and rounded once when the device field is packed (the
        legacy generation's arithmetic, full_solver.py:120,130-167); upstream's current generation evaluates them on the
        float32-rounded mesh (domain.py:392-451), a 1e-7 relative difference in ``ne``;
      * ``external_ne`` and friends work (upstream's frozen eqx.Module rejects the assignment, domain.py:310,453-461)."""

    def __init__(self, lengths, dims, *, ne_type=None, inv_brems=False, phaseshift=False, B_on=False,
                 probing_direction="z", auto_batching=True, iteration=1, region_count=1, leeway_factor=None,
                 coord_backup=None, future_dims=None, debug=False):
        self.inv_brems, self.phaseshift, self.B_on = inv_brems, phaseshift, B_on
        self.probing_direction = probing_direction
        self.ne_type = ne_type
        self.leeway_factor = 1.1 if leeway_factor is None else leeway_factor
        self.debug = debug
        # domain.py:108-132: scalar or length-3
        if np.ndim(lengths) == 0:
            lengths = [lengths] * 3
        if np.ndim(dims) == 0:
            dims = [dims] * 3
        if len(lengths) != 3:
            raise Exception("lengths must have len = 3: (x,y,z)")
        if len(dims) != 3:
            raise Exception("n must have len = 3: (x_n, y_n, z_n)")
        self.x_length, self.y_length, self.z_length = (float(v) for v in lengths)
        self.lengths = np.array([self.x_length, self.y_length, self.z_length])
        self.x_n, self.y_n, self.z_n = (int(v) for v in dims)
        self.dims = np.array([self.x_n, self.y_n, self.z_n])
        if region_count not in (None, 1):
            import warnings
            warnings.warn("synthpy_b200.ScalarDomain keeps the whole grid on one GPU (a 1024^3 packed field is 17 GB of 180 GB): "
                          f"region_count={region_count} is ignored and the domain is traced as one region "
                          "(grids beyond HBM: synthpy_b200.out_of_core.solve_out_of_core)", stacklevel=2)
        self.region_count = 1                       # no domain batching needed on a 180 GB part
        self.coord_backup = self.future_dims = None
        # domain.py:230-232: float32-rounded linspace axes
        self._axes64 = [np.linspace(-L / 2, L / 2, n) for L, n in zip(self.lengths, self.dims)]
        self.x, self.y, self.z = (np.float32(a) for a in self._axes64)
        self.ne = self.B = self.Te = self.Z = None
        self._fields = {}
        # domain.py:380-390: optional profile selected by name
        if ne_type is not None:
            getattr(self, ne_type)()

    # -- profiles (domain.py:392-451; arithmetic follows the runnable legacy code, full_solver.py:130-167)
    def _mesh(self):
        return np.meshgrid(*self._axes64, indexing="ij", copy=False)

    def _set(self, ne):
        self.ne = ne
        self._fields.clear()

    def test_null(self):
        self._set(np.zeros(tuple(self.dims)))

    def test_slab(self, s=1, ne_0=2e23):
        self._set(ne_0 * (1.0 + s * self._mesh()[0] / self.x_length))

    def test_linear_cos(self, s1=0.1, s2=0.1, ne_0=2e23, Ly=1):
        XX, YY, _ = self._mesh()
        self._set(ne_0 * (1.0 + s1 * XX / self.x_length) * (1 + s2 * np.cos(2 * np.pi * YY / Ly)))

    def test_exponential_cos(self, ne_0=1e24, Ly=1e-3, s=2e-3):
        XX, YY, _ = self._mesh()
        self._set(ne_0 * 10 ** (XX / s) * (1 + np.cos(2 * np.pi * YY / Ly)))

    def test_lens(self, ne_0=1e24, LR=1e-3):           # minimal_solver.py:192-201: Gaussian column along z
        XX, YY, _ = self._mesh()
        self._set(ne_0 * np.exp(-(np.sqrt(XX ** 2 + YY ** 2)) ** 2 / LR ** 2))

    def test_liner(self, ne_0=1e24, LR=1e-3):          # minimal_solver.py:203-212: Gaussian column along y
        XX, _, ZZ = self._mesh()
        self._set(ne_0 * np.exp(-(np.sqrt(XX ** 2 + ZZ ** 2)) ** 2 / LR ** 2))

    # -- external grids (domain.py:453-491)
    def external_ne(self, ne):
        """ne: (x_n, y_n, z_n) numpy array or CUDA torch tensor (float64 or float32), m^-3."""
        if tuple(ne.shape) != tuple(self.dims):
            raise ValueError(f"ne has shape {tuple(ne.shape)}, domain is {tuple(self.dims)}")
        self._set(ne)

    def external_B(self, B):
        self.B = B
        self._fields.clear()

    def external_Te(self, Te, Te_min=1.0):
        self.Te = np.maximum(Te_min, Te)
        self._fields.clear()

    def external_Z(self, Z):
        self.Z = Z
        self._fields.clear()

    def test_B(self, Bmax=1.0):                       # domain.py:493-503
        XX = self._mesh()[0]
        self.B = np.zeros(XX.shape + (3,))
        self.B[..., 2] = Bmax * XX / self.x_length
        self._fields.clear()

    # -- device side
    def device_field(self, lwl, *, phase=None, phase_f64=False):
        """Packed float4 {grad, n-1} grid for wavelength ``lwl`` (cached)."""
        phase = self.phaseshift if phase is None else phase
        key = (float(lwl), bool(phase), bool(phase_f64), self.probing_direction)
        if key not in self._fields:
            if self.ne is None:
                raise RuntimeError("no electron density loaded (call a test_* profile or external_ne)")
            self._fields[key] = engine.DeviceField.from_ne(
                self.ne, self.x, self.y, self.z, engine.omega_of(lwl),
                march_axis=engine.AXIS[self.probing_direction], phase=phase, phase_f64=phase_f64)
            if self.inv_brems or self.B_on:
                ne = self.ne.cpu().numpy() if hasattr(self.ne, "cpu") else self.ne
                kappa = engine.kappa_grid(ne, self.Te, self.Z, engine.omega_of(lwl)) if self.inv_brems else None
                self._fields[key].attach_channels(kappa=kappa, ne=ne if self.B_on else None, B=self.B if self.B_on else None)
        return self._fields[key]

    def release_ne(self):
        """Drop the density grid but keep the packed device fields built from it (a 1024^3 float64 grid is 8.6 GB
        that the ray kernels never read again)."""
        self.ne = None

    def cell_size(self, axis=None):
        a = engine.AXIS[self.probing_direction] if axis is None else axis
        return self.lengths[a] / (self.dims[a] - 1)

    def cleanup(self):
        pass
