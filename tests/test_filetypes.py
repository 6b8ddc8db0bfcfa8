"""VTK ImageData input/output (synthpy_b200/handle_filetypes.py; reference: src/utils/handle_filetypes.py).

vtk / pyvista are absent from the image, so the reader is held to (a) files in every layout the VTK XML format
allows, written here byte by byte independently of our writer, (b) the PVTI wrapper the reference ships
(evaluation/sergio_testing/python_cube.pvti: layout and attribute values) and the one it writes by hand, and
(c) round trips through our own writer."""
import base64
import os
import zlib

import numpy as np
import pytest

from synthpy_b200 import handle_filetypes as hf
HERE = os.path.dirname(os.path.abspath(__file__))


def _arr(shape, dtype, seed=0):
    return np.random.RandomState(seed).standard_normal(shape).astype(dtype)


@pytest.mark.parametrize("encoding,compress", [("raw", False), ("raw", True), ("base64", False), ("base64", True)])
@pytest.mark.parametrize("dtype", ["f4", "f8"])
def test_round_trip_every_layout(tmp_path, encoding, compress, dtype, capsys):
    a = _arr((6, 8, 4), dtype) * 1e24      # even sizes: the reference's spacing n // 2 is exact only then
    hf.export_pvti(a, fname=str(tmp_path / "cube"), extent_x=3e-3, extent_y=4.5e-3, extent_z=2e-3, encoding=encoding,
                   compress=compress)
    assert "succesfully saved" in capsys.readouterr().out                     # the reference prints the same two lines
    for name in ("cube.pvti", "cube.vti"):
        img, dim, spacing = hf.pvti_readin(str(tmp_path / name))
        assert dim == (6, 8, 4) and img.dtype == np.dtype(dtype)
        assert np.array_equal(img, a)
        assert np.array_equal(spacing, hf.cell_spacing(a.shape, (3e-3, 4.5e-3, 2e-3)))
        # the reference's driver rebuilds the box from these: extent = dim * spacing / 2 (pvti_trace_multiprocess.py:47-49)
        assert np.allclose(np.array(dim) * spacing / 2, [3e-3, 4.5e-3, 2e-3], rtol=1e-12)


def test_default_extents_and_large_blocks(tmp_path):
    a = _arr((40, 33, 50), "f4")                                              # > one 32 KiB compression block
    hf.export_pvti(a, fname=str(tmp_path / "d"), compress=True)
    img, dim, spacing = hf.pvti_readin(str(tmp_path / "d.pvti"))
    assert np.array_equal(img, a)
    assert np.array_equal(spacing, [1.0, 1.0, 1.0])                           # default extent n // 2 -> unit cells


def _vti(path, body, *, header_type="UInt32", byte_order="LittleEndian", compressor=None, extent="0 3 0 2 0 4",
         appended=None, app_encoding="raw"):
    comp = f' compressor="{compressor}"' if compressor else ""
    txt = (f'<?xml version="1.0"?>\n<VTKFile type="ImageData" version="0.1" byte_order="{byte_order}" header_type="{header_type}"{comp}>\n'
           f'<ImageData WholeExtent="{extent}" Origin="0 0 0" Spacing="0.5 0.25 2">\n<Piece Extent="{extent}">\n'
           f'<PointData>\n</PointData>\n<CellData Scalars="rnec">\n{body}\n</CellData>\n</Piece>\n</ImageData>\n').encode()
    if appended is not None:
        txt += f'<AppendedData encoding="{app_encoding}">\n   _'.encode() + appended + b"\n</AppendedData>\n"
    txt += b"</VTKFile>\n"
    with open(path, "wb") as fh:
        fh.write(txt)


def _check(path, a):
    img, dim, spacing = hf.pvti_readin(str(path))
    assert dim == a.shape and np.array_equal(img, a)
    assert np.array_equal(spacing, [0.5, 0.25, 2.0])


def test_hand_written_layouts(tmp_path):
    """Files assembled here from the format description, not by our writer."""
    a = _arr((3, 2, 4), "f4", 1)
    flat = a.flatten(order="F")                                               # VTK order: x fastest
    raw = flat.tobytes()
    # ascii
    _vti(tmp_path / "ascii.vti", '<DataArray type="Float32" Name="rnec" format="ascii">\n' +
         " ".join(repr(float(v)) for v in flat) + "\n</DataArray>")
    _check(tmp_path / "ascii.vti", a)
    # inline base64, UInt32 header joined with the data (current VTK) and encoded on its own (old VTK)
    h32 = np.array([len(raw)], "<u4").tobytes()
    for tag, payload in (("joined", base64.b64encode(h32 + raw)), ("split", base64.b64encode(h32) + base64.b64encode(raw))):
        _vti(tmp_path / f"inline_{tag}.vti", '<DataArray type="Float32" Name="rnec" format="binary">\n   ' + payload.decode() + "\n</DataArray>")
        _check(tmp_path / f"inline_{tag}.vti", a)
    # inline base64 + zlib in two blocks, UInt32 header [n_blocks, block, last, sizes...] as its own base64 unit
    blk = 64
    blocks = [zlib.compress(raw[i:i + blk]) for i in range(0, len(raw), blk)]
    head = np.array([len(blocks), blk, len(raw) % blk] + [len(b) for b in blocks], "<u4").tobytes()
    _vti(tmp_path / "inline_z.vti", '<DataArray type="Float32" Name="rnec" format="binary">' +
         (base64.b64encode(head) + base64.b64encode(b"".join(blocks))).decode() + "</DataArray>", compressor="vtkZLibDataCompressor")
    _check(tmp_path / "inline_z.vti", a)
    # appended raw with a second array in front (offset != 0), UInt64 header
    other = np.arange(24, dtype="<i4").tobytes()
    app = np.array([len(other)], "<u8").tobytes() + other
    off = len(app)
    app += np.array([len(raw)], "<u8").tobytes() + raw
    _vti(tmp_path / "app_raw.vti", f'<DataArray type="Float32" Name="rnec" format="appended" offset="{off}"/>\n'
         '<DataArray type="Int32" Name="id" format="appended" offset="0"/>', header_type="UInt64", appended=app)
    _check(tmp_path / "app_raw.vti", a)
    img, _, _ = hf.pvti_readin(str(tmp_path / "app_raw.vti"), array="id")
    assert np.array_equal(img, np.arange(24).reshape((3, 2, 4), order="F"))
    # appended base64 (VTK's default writer mode), offsets count encoded characters
    e0 = base64.b64encode(np.array([len(other)], "<u4").tobytes() + other)
    e1 = base64.b64encode(h32 + raw)
    _vti(tmp_path / "app_b64.vti", f'<DataArray type="Int32" Name="id" format="appended" offset="0"/>\n'
         f'<DataArray type="Float32" Name="rnec" format="appended" offset="{len(e0)}"/>', appended=e0 + e1, app_encoding="base64")
    img, _, _ = hf.pvti_readin(str(tmp_path / "app_b64.vti"), array="rnec")
    assert np.array_equal(img, a)
    # big-endian float64, appended raw
    ab = _arr((3, 2, 4), "f8", 2)
    rb = ab.flatten(order="F").astype(">f8").tobytes()
    _vti(tmp_path / "be.vti", '<DataArray type="Float64" Name="rnec" format="appended" offset="0"/>', byte_order="BigEndian",
         appended=np.array([len(rb)], ">u4").tobytes() + rb)
    _check(tmp_path / "be.vti", ab)


def test_vector_cell_data(tmp_path):
    """n_comp = 3 (e.g. a B field): (nx, ny, nz, 3) like the reference's reshape (handle_filetypes.py:111-117)."""
    v = _arr((3, 2, 4, 3), "f4", 3)
    raw = np.ascontiguousarray(v.transpose(2, 1, 0, 3)).tobytes()              # component fastest, then x, y, z
    _vti(tmp_path / "vec.vti", '<DataArray type="Float32" Name="B" NumberOfComponents="3" format="appended" offset="0"/>',
         appended=np.array([len(raw)], "<u4").tobytes() + raw)
    img, dim, _ = hf.pvti_readin(str(tmp_path / "vec.vti"))
    assert dim == (3, 2, 4, 3) and np.array_equal(img, v)


def test_reference_pvti_wrappers(tmp_path):
    # (1) the wrapper shipped with the reference (evaluation/sergio_testing/python_cube.pvti; its piece file is not
    #     shipped): same element layout and attribute values, re-typed here rather than copied
    spacing_attr = "9.900000000000001e-05 9.9e-06 9.900000000000001e-05"
    pad = " " * 24
    with open(tmp_path / "python_cube.pvti", "w") as fh:
        fh.write(f'<?xml version="1.0"?>\n{pad}<VTKFile type="PImageData" version="0.1" byte_order="LittleEndian" header_type="UInt32" '
                 f'compressor="vtkZLibDataCompressor">\n{pad}<PImageData WholeExtent="0 100 0 1000 0 100" GhostLevel="0" Origin="0 0 0" '
                 f'Spacing="{spacing_attr}">\n{pad}<PCellData Scalars="rnec">\n{pad}<PDataArray type="Float64" Name="rnec">\n'
                 f'{pad}</PDataArray>\n{pad}</PCellData>\n{pad}<Piece Extent="0 100 0 1000 0 100" Source="python_cube.vti"/>\n'
                 f'{pad}</PImageData>\n{pad}</VTKFile>')
    h = hf.pvti_header(str(tmp_path / "python_cube.pvti"))
    assert h["whole_extent"] == [0, 100, 0, 1000, 0, 100]
    assert h["cell_arrays"] == [("rnec", "Float64")]
    assert os.path.basename(h["pieces"][0][1]) == "python_cube.vti" and h["pieces"][0][0] == h["whole_extent"]
    # ... and its Spacing attribute is, digit for digit, what export_pvti's arithmetic gives for that grid
    sp = hf.cell_spacing((100, 1000, 100), (4.95e-3, 4.95e-3, 4.95e-3))
    assert f"{sp[0]!r} {sp[1]!r} {sp[2]!r}" == spacing_attr
    # (2) the wrapper text the reference writes by hand today (handle_filetypes.py:72-81): indented, declares Float32
    #     whatever the data are, compressor attribute present -- the piece file is authoritative for type and layout
    a = _arr((4, 6, 2), "f8", 4)
    hf.export_pvti(a, fname=str(tmp_path / "p"), extent_x=1.0, extent_y=1.0, extent_z=1.0)
    pad = " " * 20
    with open(tmp_path / "ref_style.pvti", "w") as fh:
        fh.write(f'<?xml version="1.0"?>\n{pad}<VTKFile type="PImageData" version="0.1" byte_order="LittleEndian" header_type="UInt32" '
                 f'compressor="vtkZLibDataCompressor">\n{pad}<PImageData WholeExtent="0 4 0 6 0 2" GhostLevel="0" Origin="0 0 0" '
                 f'Spacing="0.5 0.3333333333333333 1.0">\n{pad}<PCellData Scalars="rnec">\n{pad}<PDataArray type="Float32" Name="rnec">\n'
                 f'{pad}</PDataArray>\n{pad}</PCellData>\n{pad}<Piece Extent="0 4 0 6 0 2" Source="p.vti"/>\n{pad}</PImageData>\n{pad}</VTKFile>')
    img, dim, spacing = hf.pvti_readin(str(tmp_path / "ref_style.pvti"))
    assert np.array_equal(img, a) and img.dtype == np.float64
    assert np.array_equal(spacing, [0.5, 0.3333333333333333, 1.0])


def test_multi_piece_pvti(tmp_path):
    """A dump written by a parallel code: pieces tile the whole extent."""
    a = _arr((5, 4, 6), "f4", 5)
    cuts = [((0, 2), (0, 4), (0, 6)), ((2, 5), (0, 4), (0, 3)), ((2, 5), (0, 4), (3, 6))]
    pieces = ""
    for k, ((x0, x1), (y0, y1), (z0, z1)) in enumerate(cuts):
        sub = a[x0:x1, y0:y1, z0:z1]
        raw = sub.flatten(order="F").tobytes()
        ext = f"{x0} {x1} {y0} {y1} {z0} {z1}"
        _vti(tmp_path / f"part{k}.vti", '<DataArray type="Float32" Name="rnec" format="appended" offset="0"/>', extent=ext,
             appended=np.array([len(raw)], "<u4").tobytes() + raw)
        pieces += f'<Piece Extent="{ext}" Source="part{k}.vti"/>\n'
    with open(tmp_path / "whole.pvti", "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="PImageData" version="0.1" byte_order="LittleEndian">\n'
                 '<PImageData WholeExtent="0 5 0 4 0 6" GhostLevel="0" Origin="0 0 0" Spacing="0.5 0.25 2">\n'
                 '<PCellData Scalars="rnec"><PDataArray type="Float32" Name="rnec"/></PCellData>\n' + pieces + "</PImageData>\n</VTKFile>\n")
    _check(tmp_path / "whole.pvti", a)
    t, dim, _ = hf.pvti_readin(str(tmp_path / "whole.pvti"), device="cpu")    # the streaming (tensor) path, on the host here
    assert dim == (5, 4, 6) and t.is_contiguous() and np.array_equal(t.numpy(), a)



@pytest.mark.parametrize("encoding,compress", [("raw", False), ("base64", True)])
def test_slab_source_for_out_of_core_tracing(tmp_path, encoding, compress):
    """``pvti_slab_source``: planes [k0, k1) of the probing axis straight from the dump (what out_of_core.solve_out_of_core
    consumes) -- every direction, single- and multi-piece files, NumPy and tensor (host here) paths, and only the mapped
    byte range for a raw z-slab."""
    a = _arr((6, 8, 10), "f8", 3) * 1e24
    hf.export_pvti(a, fname=str(tmp_path / "ne"), encoding=encoding, compress=compress)
    for pd, ax in (("x", 0), ("y", 1), ("z", 2)):
        for dev in (None, "cpu"):
            src, dims, spacing = hf.pvti_slab_source(str(tmp_path / "ne.pvti"), probing_direction=pd, device=dev)
            assert dims == a.shape
            for k0, k1 in ((0, a.shape[ax]), (2, 5), (a.shape[ax] - 1, a.shape[ax])):
                idx = [slice(None)] * 3
                idx[ax] = slice(k0, k1)
                got = src(k0, k1)
                got = got if dev is None else got.numpy()
                assert got.dtype == a.dtype and np.array_equal(got, a[tuple(idx)]), (pd, dev, k0, k1)
            with pytest.raises(IndexError):
                src(3, a.shape[ax] + 1)
    src, _, _ = hf.pvti_slab_source(str(tmp_path / "ne.pvti"), scale=1e-6)
    assert np.array_equal(src(1, 4), a[:, :, 1:4] * 1e-6)
    # multi-piece dump (pieces cut in x and z): slabs that straddle the cuts
    b = _arr((5, 4, 6), "f4", 5)
    cuts = [((0, 2), (0, 4), (0, 6)), ((2, 5), (0, 4), (0, 3)), ((2, 5), (0, 4), (3, 6))]
    pieces = ""
    for k, ((x0, x1), (y0, y1), (z0, z1)) in enumerate(cuts):
        raw = b[x0:x1, y0:y1, z0:z1].flatten(order="F").tobytes()
        ext = f"{x0} {x1} {y0} {y1} {z0} {z1}"
        _vti(tmp_path / f"part{k}.vti", '<DataArray type="Float32" Name="rnec" format="appended" offset="0"/>', extent=ext,
             appended=np.array([len(raw)], "<u4").tobytes() + raw)
        pieces += f'<Piece Extent="{ext}" Source="part{k}.vti"/>\n'
    with open(tmp_path / "whole.pvti", "w") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="PImageData" version="0.1" byte_order="LittleEndian">\n'
                 '<PImageData WholeExtent="0 5 0 4 0 6" GhostLevel="0" Origin="0 0 0" Spacing="0.5 0.25 2">\n'
                 '<PCellData Scalars="rnec"><PDataArray type="Float32" Name="rnec"/></PCellData>\n' + pieces + "</PImageData>\n</VTKFile>\n")
    for pd, ax in (("x", 0), ("z", 2)):
        src, dims, _ = hf.pvti_slab_source(str(tmp_path / "whole.pvti"), probing_direction=pd)
        for k0, k1 in ((1, 4), (0, b.shape[ax]), (2, 3)):
            idx = [slice(None)] * 3
            idx[ax] = slice(k0, k1)
            assert np.array_equal(src(k0, k1), b[tuple(idx)]), (pd, k0, k1)


def test_tensor_path_equals_numpy_path_and_domain_loader(tmp_path):
    a = np.abs(_arr((8, 6, 10), "f8", 6)) * 1e24
    hf.export_pvti(a, fname=str(tmp_path / "ne"), extent_x=2e-3, extent_y=1.5e-3, extent_z=5e-3)
    t, dim, sp = hf.pvti_readin(str(tmp_path / "ne.pvti"), device="cpu")
    assert np.array_equal(t.numpy(), a) and dim == a.shape
    dom, ext = hf.domain_from_pvti(str(tmp_path / "ne.pvti"), probing_direction="y", scale=1e-6, device=None)
    assert np.allclose(ext, [2e-3, 1.5e-3, 5e-3], rtol=1e-12) and dom.probing_direction == "y"
    assert np.allclose(dom.lengths, 2 * np.array(ext)) and tuple(dom.dims) == a.shape
    assert np.array_equal(dom.ne, a * 1e-6)
    # axes as the reference's driver builds them: linspace(-extent, extent, dim) (pvti_trace_multiprocess.py:54-56)
    assert np.array_equal(dom.x, np.float32(np.linspace(-ext[0], ext[0], 8)))


def test_errors(tmp_path):
    with pytest.raises(ImportError):
        hf.hdf_readin("nothing.h5")
    with pytest.raises(Exception, match="No electron density"):
        hf.export_pvti(None, fname=str(tmp_path / "x"))
    _vti(tmp_path / "lz4.vti", '<DataArray type="Float32" Name="rnec" format="appended" offset="0"/>', compressor="vtkLZ4DataCompressor",
         appended=b"\0" * 16)
    with pytest.raises(NotImplementedError):
        hf.pvti_readin(str(tmp_path / "lz4.vti"))
    _vti(tmp_path / "nocell.vti", "")
    with pytest.raises(ValueError, match="no cell data"):
        hf.pvti_readin(str(tmp_path / "nocell.vti"))


def test_round_trip_property(tmp_path):
    """Random shapes, dtypes, layouts and extents: what is written is what is read (hypothesis)."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
    @given(shape=st.tuples(st.integers(2, 9), st.integers(2, 9), st.integers(2, 9)),
           dtype=st.sampled_from(["f4", "f8", "i4", "u2"]), encoding=st.sampled_from(["raw", "base64"]), compress=st.booleans(),
           ext=st.tuples(*[st.floats(1e-4, 10.0)] * 3), seed=st.integers(0, 2 ** 16))
    def run(shape, dtype, encoding, compress, ext, seed):
        rng = np.random.RandomState(seed)
        a = (rng.standard_normal(shape) * 1000).astype(dtype)
        base = str(tmp_path / f"h{seed}")
        hf.export_pvti(a, fname=base, extent_x=ext[0], extent_y=ext[1], extent_z=ext[2], encoding=encoding, compress=compress)
        img, dim, spacing = hf.pvti_readin(base + ".pvti")
        assert dim == shape and img.dtype == a.dtype and np.array_equal(img, a)
        assert np.array_equal(spacing, hf.cell_spacing(shape, ext))
        t, _, _ = hf.pvti_readin(base + ".vti", device="cpu")
        assert t.is_contiguous() and np.array_equal(t.numpy(), a)
    run()
