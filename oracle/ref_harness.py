"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Runs the UNMODIFIED reference modules from /root/reference on the CPU so that
golden vectors can be generated (oracle/make_goldens.py) and the numpy
restatement (oracle/modal_oracle.py) can be pinned against the real thing.

The reference needs three third-party pieces that are absent in this image;
each is replaced by a minimal stand-in injected through ``sys.modules``:

* ``torch_scatter.scatter`` (reference use: src/diffelastic/deform.py:165 'sum',
  src/diffelastic/mesh.py:176 'min')
* ``meshio.read`` for gmsh-2.2 binary ``.msh`` (reference use:
  src/diffelastic/mesh.py:48)
* ``Tensor.cuda`` / ``Module.cuda`` -> identity (the reference hard-codes
  ``.cuda()``; this container has no GPU)

This file only exists where /root/reference exists (the build container).  The
GPU box never sees the reference; tests there use tests/golden/*.npz.
"""
import os
import struct
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("DIFFSOUND_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "diffelastic"))


# ----------------------------------------------------------------------------
# shims
# ----------------------------------------------------------------------------
def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    if dim < 0:
        dim += src.dim()
    if index.dim() == 1 and src.dim() > 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.reshape(shape).expand_as(src)
    if dim_size is None:
        dim_size = int(index.max()) + 1
    shape = list(src.shape)
    shape[dim] = dim_size
    if reduce in ("sum", "add"):
        res = torch.zeros(shape, dtype=src.dtype, device=src.device)
        return res.scatter_add_(dim, index, src)
    if reduce == "min":
        res = torch.zeros(shape, dtype=src.dtype, device=src.device)
        return res.scatter_reduce_(dim, index, src, "amin", include_self=False)
    raise NotImplementedError(reduce)


class _Cell:
    def __init__(self, type_, data):
        self.type = type_
        self.data = data


class _Mesh:
    def __init__(self, points, cells):
        self.points = points
        self.cells = cells
        self.cells_dict = {c.type: c.data for c in cells}


_GMSH_TYPES = {1: ("line", 2), 2: ("triangle", 3), 4: ("tetra", 4), 15: ("vertex", 1)}


def read_gmsh22(filename):
    """Parse a gmsh 2.2 .msh file (binary or ascii); returns points (f64) and
    a dict type -> connectivity (0-based)."""
    with open(filename, "rb") as f:
        buf = f.read()
    pos = 0

    def line():
        nonlocal pos
        e = buf.index(b"\n", pos)
        s = buf[pos:e].decode()
        pos = e + 1
        return s.strip()

    assert line() == "$MeshFormat"
    ver, ftype, dsize = line().split()
    binary = int(ftype) == 1
    if binary:
        one = struct.unpack("i", buf[pos:pos + 4])[0]
        assert one == 1, "endianness"
        pos += 4
        if buf[pos:pos + 1] == b"\n":
            pos += 1
    assert line() == "$EndMeshFormat"
    points = None
    cells = {}
    while pos < len(buf):
        tag = line()
        if tag == "$Nodes":
            n = int(line())
            if binary:
                rec = np.dtype([("id", "<i4"), ("x", "<f8", (3,))])
                arr = np.frombuffer(buf, dtype=rec, count=n, offset=pos)
                pos += n * rec.itemsize
                if buf[pos:pos + 1] == b"\n":
                    pos += 1
                ids = arr["id"].astype(np.int64)
                pts = arr["x"].copy()
            else:
                rows = [line().split() for _ in range(n)]
                ids = np.array([int(r[0]) for r in rows])
                pts = np.array([[float(v) for v in r[1:4]] for r in rows])
            assert np.array_equal(ids, np.arange(1, n + 1))
            points = pts
            assert line() == "$EndNodes"
        elif tag == "$Elements":
            total = int(line())
            got = 0
            if binary:
                while got < total:
                    etype, cnt, ntags = struct.unpack("<3i", buf[pos:pos + 12])
                    pos += 12
                    name, nn = _GMSH_TYPES[etype]
                    w = 1 + ntags + nn
                    arr = np.frombuffer(buf, dtype="<i4", count=cnt * w, offset=pos).reshape(cnt, w)
                    pos += cnt * w * 4
                    cells.setdefault(name, []).append(arr[:, 1 + ntags:].astype(np.int64) - 1)
                    got += cnt
                if buf[pos:pos + 1] == b"\n":
                    pos += 1
            else:
                for _ in range(total):
                    r = [int(v) for v in line().split()]
                    name, nn = _GMSH_TYPES[r[1]]
                    cells.setdefault(name, []).append(np.array([r[3 + r[2]:]], dtype=np.int64) - 1)
                    got += 1
            assert line() == "$EndElements"
        else:
            # skip unknown (possibly binary) section
            end = ("$End" + tag[1:]).encode()
            at = buf.find(end, pos)
            if at < 0:
                break
            pos = at + len(end) + 1
    return points, {k: np.concatenate(v, axis=0) for k, v in cells.items()}


def _meshio_read(filename):
    points, cells = read_gmsh22(filename)
    order = [k for k in ("tetra", "triangle", "line", "vertex") if k in cells]
    return _Mesh(points, [_Cell(k, cells[k]) for k in order])


_installed = False


def install_shims():
    global _installed
    if _installed:
        return
    ts = types.ModuleType("torch_scatter")
    ts.scatter = _scatter
    sys.modules.setdefault("torch_scatter", ts)
    mi = types.ModuleType("meshio")
    mi.read = _meshio_read
    sys.modules.setdefault("meshio", mi)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def ref_modules():
    """Import and return the reference's hot-path modules (unmodified)."""
    install_shims()
    import importlib
    out = types.SimpleNamespace()
    out.diff_model = importlib.import_module("src.diffelastic.diff_model")
    out.mesh = importlib.import_module("src.diffelastic.mesh")
    out.deform = importlib.import_module("src.diffelastic.deform")
    out.gauss = importlib.import_module("src.diffelastic.gauss")
    out.shape_func = importlib.import_module("src.diffelastic.shape_func")
    out.mass_matrix = importlib.import_module("src.diffelastic.mass_matrix")
    out.material_model = importlib.import_module("src.diffelastic.material_model")
    out.oscillator = importlib.import_module("src.ddsp.oscillator")
    out.lobpcg = importlib.import_module("src.lobpcg")
    return out


def ref_path(*parts):
    return os.path.join(REF_ROOT, *parts)
