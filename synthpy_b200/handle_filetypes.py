"""VTK ImageData (``.vti`` / ``.pvti``) input and output with the call shape of the reference's
``src/utils/handle_filetypes.py`` -- the data format on the input side of the ray path.

The reference reads with ``vtk.vtkXMLPImageDataReader`` and writes with ``pyvista`` (handle_filetypes.py:11-122);
neither is needed here: the VTK XML ImageData format is restated directly (header in XML, arrays inline or in
one appended section; ascii, base64 or raw; optionally zlib-compressed in blocks; UInt32 or UInt64 size headers).
**Parity unpinned** against vtk itself (vtk / pyvista are not installed in the build image); what is pinned:
the PVTI wrapper the reference writes by hand (handle_filetypes.py:72-81) and the one the reference ships
(evaluation/sergio_testing/python_cube.pvti, its layout and attribute values in tests/test_filetypes.py) parse, and the array
semantics of ``pvti_readin`` (first cell array, Fortran-order reshape to (nx, ny, nz), spacing) are the reference's.

``pvti_readin(..., device='cuda')`` streams the array into HBM piece by piece through pinned memory and returns a
CUDA tensor that ``ScalarDomain.external_ne`` takes as is; a raw appended section is memory-mapped, so a 1024^3
grid never exists twice on the host.
"""
import base64
import os
import re
import xml.etree.ElementTree as ET
import zlib

import numpy as np

_VTK_TYPES = {"Float32": "f4", "Float64": "f8", "Int8": "i1", "UInt8": "u1", "Int16": "i2", "UInt16": "u2",
              "Int32": "i4", "UInt32": "u4", "Int64": "i8", "UInt64": "u8"}
_NP_TO_VTK = {np.dtype(v).str[1:]: k for k, v in _VTK_TYPES.items()}


# ------------------------------------------------------------------------------------------------ reading
def _b64_chars(nbytes):
    return 4 * ((nbytes + 2) // 3)


class _VTIFile:
    """One serial ``.vti`` file: parsed header plus lazy access to its data arrays."""

    def __init__(self, path):
        self.path = path
        # the appended section is not XML (bytes after '_'): parse only what precedes it, never read it whole
        raw, m = b"", None
        with open(path, "rb") as fh:
            while m is None:
                more = fh.read(1 << 20)
                raw += more
                m = re.search(rb"<AppendedData[^>]*>", raw)
                if not more:
                    break
            while m is not None and b"_" not in raw[m.end():]:
                more = fh.read(4096)
                if not more:
                    raise ValueError(f"{path}: AppendedData without '_' marker")
                raw += more
        if m is not None:
            self._app_off = raw.index(b"_", m.end()) + 1
            enc = re.search(rb'encoding\s*=\s*"(\w+)"', m.group(0))
            self._app_enc = enc.group(1).decode() if enc else "base64"
            xml = raw[:m.end()] + b"</AppendedData></VTKFile>"
        else:
            xml = raw
        root = ET.fromstring(xml)
        if root.tag != "VTKFile" or root.get("type") != "ImageData":
            raise ValueError(f"{path}: not a VTK ImageData file (type={root.get('type')!r})")
        self.bo = "<" if root.get("byte_order", "LittleEndian") == "LittleEndian" else ">"
        self.hdr = np.dtype(self.bo + _VTK_TYPES[root.get("header_type", "UInt32")])
        comp = root.get("compressor")
        if comp not in (None, "", "vtkZLibDataCompressor"):
            raise NotImplementedError(f"{path}: compressor {comp} (only vtkZLibDataCompressor is supported)")
        self.compressed = bool(comp)
        img = root.find("ImageData")
        self.whole_extent = [int(v) for v in img.get("WholeExtent").split()]
        self.origin = np.array([float(v) for v in img.get("Origin", "0 0 0").split()])
        self.spacing = np.array([float(v) for v in img.get("Spacing", "1 1 1").split()])
        self.pieces = []
        for p in img.findall("Piece"):
            ext = [int(v) for v in p.get("Extent").split()]
            cd, pd = p.find("CellData"), p.find("PointData")
            self.pieces.append({"extent": ext,
                                "cell": [] if cd is None else cd.findall("DataArray"),
                                "point": [] if pd is None else pd.findall("DataArray")})

    # -- one DataArray element -> flat numpy array (file order: x fastest)
    def _decode_blocks(self, get, dtype):
        """``get(offset, nbytes)`` returns decoded bytes of the array's byte stream (header first)."""
        hs = self.hdr.itemsize
        if not self.compressed:
            n = int(np.frombuffer(get(0, hs), self.hdr)[0])
            return np.frombuffer(get(hs, n), dtype)
        nb, bs, last = (int(v) for v in np.frombuffer(get(0, 3 * hs), self.hdr))
        sizes = np.frombuffer(get(3 * hs, nb * hs), self.hdr).astype(np.int64)
        total = (nb - 1) * bs + (last if last else bs) if nb else 0
        out = np.empty(total, np.uint8)
        off, pos = (3 + nb) * hs, 0
        for s in sizes:
            blk = zlib.decompress(get(off, int(s)))
            out[pos:pos + len(blk)] = np.frombuffer(blk, np.uint8)
            off += int(s)
            pos += len(blk)
        return out[:pos].view(dtype)

    def _from_base64(self, text, dtype):
        """Inline or appended base64: the size header is a base64 unit of its own when the data are compressed,
        and (depending on the writer's version) either separate or joined when they are not."""
        hs = self.hdr.itemsize
        if self.compressed:
            nb = int(np.frombuffer(base64.b64decode(text[:_b64_chars(3 * hs)])[:hs], self.hdr)[0])
            hchars = _b64_chars((3 + nb) * hs)
            head = base64.b64decode(text[:hchars])
            sizes = np.frombuffer(head[3 * hs:(3 + nb) * hs], self.hdr).astype(np.int64)
            body = base64.b64decode(text[hchars:hchars + _b64_chars(int(sizes.sum()))])
            stream = head[:(3 + nb) * hs] + body
            return self._decode_blocks(lambda o, n: stream[o:o + n], dtype)
        n = int(np.frombuffer(base64.b64decode(text[:_b64_chars(hs)])[:hs], self.hdr)[0])
        joined = base64.b64decode(text[:_b64_chars(hs + n)])
        if len(joined) >= hs + n:
            return np.frombuffer(joined[hs:hs + n], dtype)
        hchars = _b64_chars(hs)                      # header encoded on its own, data follow as a second unit
        return np.frombuffer(base64.b64decode(text[hchars:hchars + _b64_chars(n)])[:n], dtype)

    def array(self, elem):
        dtype = np.dtype(self.bo + _VTK_TYPES[elem.get("type")])
        fmt = elem.get("format", "ascii")
        if fmt == "ascii":
            return np.array(elem.text.split(), dtype=dtype.newbyteorder("="))
        if fmt == "binary":
            return self._from_base64("".join(elem.text.split()).encode(), dtype)
        if fmt != "appended":
            raise ValueError(f"{self.path}: DataArray format {fmt!r}")
        off = int(elem.get("offset", "0"))
        if self._app_enc == "base64":
            return self._from_base64(np.memmap(self.path, np.uint8, "r", offset=self._app_off + off), dtype)
        base = self._app_off + off
        if not self.compressed:                      # raw + uncompressed: map the file, nothing is copied here
            with open(self.path, "rb") as fh:
                fh.seek(base)
                n = int(np.frombuffer(fh.read(self.hdr.itemsize), self.hdr)[0])
            return np.memmap(self.path, dtype=dtype, mode="r", offset=base + self.hdr.itemsize, shape=(n // dtype.itemsize,))
        with open(self.path, "rb") as fh:
            def get(o, n):
                fh.seek(base + o)
                return fh.read(n)
            return self._decode_blocks(get, dtype)


def _cells(ext):
    return tuple(ext[2 * a + 1] - ext[2 * a] for a in range(3))


def pvti_header(filename):
    """Parses a ``.pvti`` wrapper without touching its pieces: whole extent (6 ints), spacing, origin, declared cell
    arrays [(name, vtk type)] and pieces [(extent, source path)]."""
    root = ET.parse(filename).getroot()
    if root.tag != "VTKFile" or root.get("type") != "PImageData":
        raise ValueError(f"{filename}: not a VTK PImageData file")
    pimg = root.find("PImageData")
    here = os.path.dirname(os.path.abspath(filename))
    pcd = pimg.find("PCellData")
    return {"whole_extent": [int(v) for v in pimg.get("WholeExtent").split()],
            "spacing": np.array([float(v) for v in pimg.get("Spacing", "1 1 1").split()]),
            "origin": np.array([float(v) for v in pimg.get("Origin", "0 0 0").split()]),
            "cell_arrays": [] if pcd is None else [(e.get("Name"), e.get("type")) for e in pcd.findall("PDataArray")],
            "pieces": [([int(x) for x in p.get("Extent").split()], os.path.join(here, p.get("Source")))
                       for p in pimg.findall("Piece")]}


def _pieces_of(filename):
    """[(vti, piece)] of a serial or parallel ImageData file, plus whole extent and spacing."""
    with open(filename, "rb") as fh:
        head = fh.read(4096)
    if b"PImageData" not in head:
        v = _VTIFile(filename)
        return [(v, p) for p in v.pieces], v.whole_extent, v.spacing
    h = pvti_header(filename)
    whole, spacing = h["whole_extent"], h["spacing"]
    out = []
    for ext, src in h["pieces"]:
        v = _VTIFile(src)
        for piece in v.pieces:
            if len(v.pieces) == 1:                   # the P-file's extent is authoritative for a single-piece source
                piece = dict(piece, extent=ext)
            out.append((v, piece))
    return out, whole, spacing


def pvti_readin(filename, *, array=0, device=None):
    """Reads the first cell-data array of a ``.pvti`` (or ``.vti``) file: the electron density of a simulation dump.

    Returns ``(img, img.shape, spacing)`` exactly like the reference (handle_filetypes.py:89-122): ``img`` is
    (nx, ny, nz) -- the file stores x fastest, i.e. the Fortran-order reshape the reference applies -- or
    (nx, ny, nz, n_comp) for vectors, ``spacing`` the three cell sizes.  ``array`` selects another cell array by
    index or name.  With ``device`` ('cuda' / torch.device) ``img`` is a C-contiguous torch tensor on that device,
    uploaded piece by piece through pinned memory."""
    pieces, whole, spacing = _pieces_of(filename)
    if not pieces:
        raise ValueError(f"{filename}: no pieces")
    dims = _cells(whole)
    lo = [whole[0], whole[2], whole[4]]

    def pick(piece, vti):
        arrs = piece["cell"]
        if not arrs:
            raise ValueError(f"{vti.path}: no cell data (the reference reads GetCellData().GetArray(0))")
        if isinstance(array, str):
            for e in arrs:
                if e.get("Name") == array:
                    return e
            raise KeyError(array)
        return arrs[array]

    first = pick(pieces[0][1], pieces[0][0])
    n_comp = int(first.get("NumberOfComponents", "1"))
    dtype = np.dtype(_VTK_TYPES[first.get("type")])
    tail = (n_comp,) if n_comp > 1 else ()

    if device is None:
        if len(pieces) == 1 and _cells(pieces[0][1]["extent"]) == dims:
            flat = pieces[0][0].array(first)
            img = np.asarray(flat).astype(dtype, copy=False).reshape(dims[::-1] + tail)
            img = img.transpose(2, 1, 0, *([3] if tail else []))          # == v.reshape(vec, order='F')
            return img, img.shape, spacing
        img = np.empty(dims + tail, dtype, order="F" if not tail else "C")
        for vti, piece in pieces:
            e = piece["extent"]
            c = _cells(e)
            blk = np.asarray(vti.array(pick(piece, vti))).reshape(c[::-1] + tail).transpose(2, 1, 0, *([3] if tail else []))
            img[e[0] - lo[0]:e[1] - lo[0], e[2] - lo[1]:e[3] - lo[1], e[4] - lo[2]:e[5] - lo[2]] = blk
        return img, img.shape, spacing

    import torch
    dev = torch.device(device)
    tdtype = torch.from_numpy(np.empty(0, dtype)).dtype
    img = torch.empty(dims + tail, dtype=tdtype, device=dev)
    chunk = 1 << 26                                                        # elements per pinned staging buffer
    pin = torch.cuda.is_available() and dev.type == "cuda"
    stage = [torch.empty(chunk, dtype=tdtype, pin_memory=pin) for _ in range(2)]
    done = [None, None]
    for vti, piece in pieces:
        e = piece["extent"]
        c = _cells(e)
        flat = vti.array(pick(piece, vti))
        dst = torch.empty(flat.shape[0], dtype=tdtype, device=dev)
        for k, s in enumerate(range(0, flat.shape[0], chunk)):
            b = k & 1
            if done[b] is not None:
                done[b].synchronize()
            n = min(chunk, flat.shape[0] - s)
            stage[b][:n].numpy()[:] = flat[s:s + n]                       # page cache / decoded buffer -> pinned
            dst[s:s + n].copy_(stage[b][:n], non_blocking=True)
            if pin:
                done[b] = torch.cuda.Event()
                done[b].record()
        blk = dst.view(c[::-1] + tail).permute(2, 1, 0, *([3] if tail else []))
        img[e[0] - lo[0]:e[1] - lo[0], e[2] - lo[1]:e[3] - lo[1], e[4] - lo[2]:e[5] - lo[2]] = blk
        del dst
    return img, tuple(img.shape), spacing


vti_readin = pvti_readin


def pvti_slab_source(filename, *, probing_direction="z", array=0, device=None, scale=1.0):
    """Out-of-core access to the first cell array of a ``.pvti`` / ``.vti`` dump: returns ``(source, dims, spacing)`` where
    ``source(k0, k1)`` is the sub-grid of planes [k0, k1) of the probing axis, shaped like ``ne[..., k0:k1]`` (or the
    corresponding slice for 'x' / 'y') -- what ``out_of_core.solve_out_of_core`` asks for, slab by slab.

    The file stores x fastest and z slowest, so for 'z' a slab of a raw, uncompressed piece is one contiguous byte range of
    the memory-mapped file: only those pages are read.  Other encodings decode the piece once and keep it.  With ``device``
    the planes are uploaded in file order and transposed to (x, y, z) on the device; without it a NumPy array is returned."""
    pieces, whole, spacing = _pieces_of(filename)
    if not pieces:
        raise ValueError(f"{filename}: no pieces")
    p = {"x": 0, "y": 1, "z": 2}[probing_direction]
    dims = _cells(whole)
    lo = [whole[0], whole[2], whole[4]]
    cache = {}

    def flat_of(i):
        if i not in cache:
            vti, piece = pieces[i]
            arrs = piece["cell"]
            if not arrs:
                raise ValueError(f"{vti.path}: no cell data")
            elem = arrs[array] if not isinstance(array, str) else next(e for e in arrs if e.get("Name") == array)
            if int(elem.get("NumberOfComponents", "1")) != 1:
                raise ValueError("pvti_slab_source reads scalar cell arrays")
            cache[i] = vti.array(elem)
        return cache[i]

    first = pieces[0][1]["cell"]
    if not first:
        raise ValueError(f"{pieces[0][0].path}: no cell data")
    dtype = np.dtype(_VTK_TYPES[(first[array] if not isinstance(array, str) else next(e for e in first if e.get("Name") == array)).get("type")])

    def source(k0, k1):
        k0, k1 = int(k0), int(k1)
        if not 0 <= k0 < k1 <= dims[p]:
            raise IndexError(f"planes [{k0}, {k1}) outside the grid's {dims[p]}")
        shape = list(dims)
        shape[p] = k1 - k0
        if device is None:
            out = np.empty(shape, dtype)
        else:
            import torch
            out = torch.empty(shape, dtype=torch.from_numpy(np.empty(0, dtype)).dtype, device=device)
        for i, (vti, piece) in enumerate(pieces):
            e = piece["extent"]
            e_lo, e_hi = e[2 * p] - lo[p], e[2 * p + 1] - lo[p]
            a, b = max(k0, e_lo), min(k1, e_hi)
            if a >= b:
                continue
            c = _cells(e)
            blk = np.asarray(flat_of(i)).reshape(c[::-1])                  # (z, y, x) of this piece, a view
            idx = [slice(None)] * 3
            idx[2 - p] = slice(a - e_lo, b - e_lo)
            sub = blk[tuple(idx)]
            dst = [slice(e[0] - lo[0], e[1] - lo[0]), slice(e[2] - lo[1], e[3] - lo[1]), slice(e[4] - lo[2], e[5] - lo[2])]
            dst[p] = slice(a - k0, b - k0)
            if device is None:
                out[tuple(dst)] = sub.transpose(2, 1, 0)
            else:
                import torch
                import warnings
                with warnings.catch_warnings():              # a mapped file range is read-only; it is only read from
                    warnings.simplefilter("ignore", UserWarning)
                    host = torch.from_numpy(np.ascontiguousarray(sub).astype(dtype.newbyteorder("="), copy=False))
                out[tuple(dst)] = host.to(device).permute(2, 1, 0)
        if scale != 1.0:
            out = out * scale
        return out

    return source, dims, spacing


# ------------------------------------------------------------------------------------------------ writing
def _write_vti(path, arr_f, cells, spacing, name, encoding, compress, block=1 << 15):
    """One-piece ImageData file with ``arr_f`` (flat, x fastest) as cell data in an appended section."""
    vtk_type = _NP_TO_VTK[arr_f.dtype.str[1:]]
    ext = f"0 {cells[0]} 0 {cells[1]} 0 {cells[2]}"
    sp = f"{spacing[0]!r} {spacing[1]!r} {spacing[2]!r}"
    head = (f'<?xml version="1.0"?>\n<VTKFile type="ImageData" version="1.0" byte_order="LittleEndian" header_type="UInt64"'
            + (' compressor="vtkZLibDataCompressor"' if compress else "") + ">\n"
            f'  <ImageData WholeExtent="{ext}" Origin="0 0 0" Spacing="{sp}">\n'
            f'    <Piece Extent="{ext}">\n      <PointData/>\n      <CellData Scalars="{name}">\n'
            f'        <DataArray type="{vtk_type}" Name="{name}" format="appended" offset="0"/>\n'
            f"      </CellData>\n    </Piece>\n  </ImageData>\n"
            f'  <AppendedData encoding="{encoding}">\n   _')
    data = arr_f.astype(arr_f.dtype.newbyteorder("<"), copy=False)
    with open(path, "wb") as fh:
        fh.write(head.encode())
        if compress:
            raw = data.tobytes()
            blocks = [zlib.compress(raw[i:i + block]) for i in range(0, len(raw), block)]
            last = len(raw) % block
            hdr = np.array([len(blocks), block, last] + [len(b) for b in blocks], "<u8").tobytes()
            if encoding == "base64":
                fh.write(base64.b64encode(hdr) + base64.b64encode(b"".join(blocks)))
            else:
                fh.write(hdr)
                for b in blocks:
                    fh.write(b)
        else:
            hdr = np.array([data.nbytes], "<u8").tobytes()
            if encoding == "base64":
                fh.write(base64.b64encode(hdr + data.tobytes()))
            else:
                fh.write(hdr)
                data.tofile(fh)
        fh.write(b"\n  </AppendedData>\n</VTKFile>\n")


def cell_spacing(shape, extents):
    """handle_filetypes.py:48-58: cell size = max(linspace(-e, e, n)) / (n // 2) per axis."""
    return [float(np.max(np.linspace(-e, e, n)) / (n // 2)) for e, n in zip(extents, shape)]


def export_pvti(arr, fname=None, extent_x=None, extent_y=None, extent_z=None, *, name="rnec", encoding="raw",
                compress=False):
    """Writes a 3-D array as ``{fname}.vti`` + ``{fname}.pvti`` (cell data ``rnec``), as the reference does
    (handle_filetypes.py:11-87): grid of shape+1 points, origin 0, spacing ``extent / (n // 2)`` per axis
    (default extents ``n // 2``), array flattened in Fortran order, PVTI wrapper with one piece.
    ``encoding`` 'raw' | 'base64' and ``compress`` choose the layout of the appended section (the reference's
    writer, pyvista, uses VTK's default); numpy arrays or torch tensors (any device) are accepted."""
    if fname is None:                                                     # handle_filetypes.py:19-27
        import datetime as dt
        now = dt.datetime.now()
        fname = f"./plasma_PVTI_{now.day}_{now.month}_{now.year}_{now.hour}_{now.minute}"
    if hasattr(arr, "detach"):
        arr = arr.detach().cpu().numpy()
    try:
        arr = np.asarray(arr)
        shape = arr.shape
        assert len(shape) == 3
    except Exception:
        raise Exception("No electron density currently loaded!")
    if arr.dtype.str[1:] not in _NP_TO_VTK:
        raise TypeError(f"dtype {arr.dtype} has no VTK type")
    if encoding not in ("raw", "base64"):
        raise ValueError("encoding must be 'raw' or 'base64'")
    ext = [shape[a] // 2 if e is None else e for a, e in enumerate((extent_x, extent_y, extent_z))]
    spacing = cell_spacing(shape, ext)
    _write_vti(f"{fname}.vti", arr.flatten(order="F"), shape, spacing, name, encoding, compress)
    print(f"VTI saved under {fname}.vti")
    rel = fname.split("/")[-1]
    wext = f"0 {shape[0]} 0 {shape[1]} 0 {shape[2]}"
    content = (f'<?xml version="1.0"?>\n<VTKFile type="PImageData" version="0.1" byte_order="LittleEndian" header_type="UInt64"'
               + (' compressor="vtkZLibDataCompressor"' if compress else "") + ">\n"
               f'<PImageData WholeExtent="{wext}" GhostLevel="0" Origin="0 0 0" Spacing="{spacing[0]!r} {spacing[1]!r} {spacing[2]!r}">\n'
               f'  <PCellData Scalars="{name}">\n    <PDataArray type="{_NP_TO_VTK[arr.dtype.str[1:]]}" Name="{name}">\n'
               f"    </PDataArray>\n  </PCellData>\n"
               f'  <Piece Extent="{wext}" Source="{rel}.vti"/>\n</PImageData>\n</VTKFile>\n')
    with open(f"{fname}.pvti", "w") as fh:
        fh.write(content)
    print(f"Scalar Domain electron density succesfully saved under {fname}.pvti !")


def hdf_readin(filename):
    """FLASH HDF5 -> ne covering grid (handle_filetypes.py:124-150).  Upstream does this with ``yt`` (AMR covering
    grid at the finest level); yt / h5py are not in this image and re-gridding AMR data is outside the ray path."""
    raise ImportError("hdf_readin needs yt (not available here); convert the dump with the reference's hdf_to_pvti "
                      "and load the .pvti with pvti_readin")


def hdf_to_pvti(hdf_filename, pvti_filename):                             # handle_filetypes.py:152-161
    ne, dims, spacing = hdf_readin(hdf_filename)
    export_pvti(ne, fname=pvti_filename, extent_x=dims[0] * spacing[0] / 2, extent_y=dims[1] * spacing[1] / 2,
                extent_z=dims[2] * spacing[2] / 2)


def domain_from_pvti(filename, *, probing_direction="z", scale=1.0, device="cuda", **domain_kw):
    """The reference drivers' ``calculate_field`` (examples/jobs/run_scripts/pvti_trace_multiprocess.py:45-65):
    read ne, centre the box on the origin with half-lengths ``dim * spacing / 2``, load it into a ScalarDomain
    (``scale`` is the unit factor the drivers apply, e.g. 1e12 there).  Returns (domain, (extent_x, extent_y, extent_z))."""
    from .domain import ScalarDomain
    ne, dim, spacing = pvti_readin(filename, device=device)
    ext = tuple(float(dim[a] * spacing[a] / 2) for a in range(3))
    dom = ScalarDomain([2 * e for e in ext], list(dim[:3]), probing_direction=probing_direction, **domain_kw)
    dom.external_ne(ne * scale if scale != 1.0 else ne)
    return dom, ext
