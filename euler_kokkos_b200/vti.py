"""Reader/writer helpers for the reference's VTK ImageData (.vti) dumps.

Format written by the reference's save_VTK_{2D,3D} (src/utils/io/IO_VTK.cpp:213-410) in its default
binary mode: XML header, then ``<AppendedData encoding="raw">_`` followed, per variable in index
order, by a UInt64 byte count and nx*ny*nz little-endian Float64 values (i fastest), interior cells
only.  The C++ host driver (csrc/host/IO_VTK.cpp) writes the same layout; this module is the
test/bench-side reader.
"""
from __future__ import annotations

import re

import numpy as np


def read_vti(path: str):
    """Return (names, data) with data shaped (nvar, nz, ny, nx), Float64."""
    with open(path, "rb") as f:
        blob = f.read()
    marker = blob.find(b"<AppendedData")
    if marker < 0:
        raise ValueError(f"{path}: not a raw-appended .vti (ascii mode is not supported)")
    header = blob[:marker].decode("ascii", errors="replace")
    m = re.search(r'WholeExtent="\s*(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)', header)
    if not m:
        raise ValueError(f"{path}: no WholeExtent")
    e = [int(x) for x in m.groups()]
    nx, ny, nz = e[1] - e[0], e[3] - e[2], e[5] - e[4]
    nz = max(nz, 1)
    arrays = re.findall(r'<DataArray type="(\w+)" Name="(\w+)" format="appended" offset="(\d+)"', header)
    start = blob.find(b"_", marker) + 1
    names, out = [], []
    n = nx * ny * nz
    for typ, name, off in arrays:
        if typ != "Float64":
            raise ValueError(f"{path}: array {name} is {typ}, expected Float64")
        o = start + int(off)
        nbytes = int(np.frombuffer(blob, dtype="<u8", count=1, offset=o)[0])
        if nbytes != 8 * n:
            raise ValueError(f"{path}: array {name} has {nbytes} bytes, expected {8 * n}")
        a = np.frombuffer(blob, dtype="<f8", count=n, offset=o + 8).reshape(nz, ny, nx)
        names.append(name)
        out.append(a)
    return names, np.stack(out).astype(np.float64)


def read_pvti(path: str):
    """Assemble a decomposed dump: parse the .pvti header (write_pvti_header, IO_VTK.cpp:857-1011), read every piece
    and place it by its extent.  Returns (names, data) with data shaped (nvar, NZ, NY, NX) for the global box."""
    import os

    with open(path) as f:
        text = f.read()
    m = re.search(r'<PImageData WholeExtent="\s*(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(\d+)', text)
    if not m:
        raise ValueError(f"{path}: no PImageData WholeExtent")
    e = [int(x) for x in m.groups()]
    NX, NY, NZ = e[1] - e[0], e[3] - e[2], max(e[5] - e[4], 1)
    names = re.findall(r'<PDataArray type="Float64" Name="(\w+)"', text)
    out = np.full((len(names), NZ, NY, NX), np.nan)
    for pm in re.finditer(r'<Piece Extent="([^"]+)" Source="([^"]+)"', text):
        ext = [int(x) for x in pm.group(1).split()]
        pnames, data = read_vti(os.path.join(os.path.dirname(path), pm.group(2)))
        assert pnames == names, (pnames, names)
        z0 = ext[4] if NZ > 1 else 0
        out[:, z0:z0 + data.shape[1], ext[2]:ext[2] + data.shape[2], ext[0]:ext[0] + data.shape[3]] = data
    return names, out


def embed_interior(interior: np.ndarray, ghost: int, dim: int) -> np.ndarray:
    """(nvar,nz,ny,nx) interior -> ghosted SoA (nvar,ksize,jsize,isize); ghosts are zero-filled."""
    nv, nz, ny, nx = interior.shape
    g = ghost
    if dim == 3:
        U = np.zeros((nv, nz + 2 * g, ny + 2 * g, nx + 2 * g))
        U[:, g:-g, g:-g, g:-g] = interior
    else:
        U = np.zeros((nv, 1, ny + 2 * g, nx + 2 * g))
        U[:, :, g:-g, g:-g] = interior
    return U


def interior_of(U: np.ndarray, ghost: int, dim: int) -> np.ndarray:
    g = ghost
    if dim == 3:
        return U[:, g:-g, g:-g, g:-g]
    return U[:, :, g:-g, g:-g]
