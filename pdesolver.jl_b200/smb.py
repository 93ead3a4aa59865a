"""Reader for PUMI ``.smb`` mesh parts: the reference's fixtures and benchmark meshes
(``src/mesh_files/*.smb``, loaded there through PumiInterface.jl, which is not vendored).

Layout (decoded from the files): big-endian; a 48-byte header ``magic, version, dim, nparts`` followed
by the entity counts ``vertex, edge, triangle, quad, hex, prism, pyramid, tet``; then the downward adjacencies of
every entity type above the vertices, in that order and 0-based (edge -> 2 vertices, triangle -> 3 edges,
tet -> 4 triangles); then ``nvertex x 3`` float64 coordinates.  Classification, parametric coordinates, remote
copies and tags follow and are not needed for the element/vertex description that ``mesh.simplex_mesh`` consumes
(it rebuilds interfaces, boundary faces, permutations and metrics itself, so every result that does not depend on
PUMI's element numbering is reproduced: error norms, convergence rates, residual norms).
"""
from __future__ import annotations

import struct

import numpy as np

_HEADER = struct.Struct(">12I")


def read_smb(path):
    """Returns ``(vertex_coords[nV, dim], simplices[nE, dim+1], dim)`` of a triangle (2D) or tet (3D) mesh part."""
    with open(path, "rb") as f:
        d = f.read()
    magic, version, dim, nparts, nv, ne, nt, nquad, nhex, nprism, npyr, ntet = _HEADER.unpack_from(d, 0)
    if magic != 0 or version not in (4, 5, 6):
        raise ValueError(f"{path}: not a version 4-6 .smb file (magic {magic}, version {version})")
    if nquad or nhex or nprism or npyr:
        raise ValueError(f"{path}: only simplex meshes are supported")
    off = _HEADER.size
    edges = np.frombuffer(d, dtype=">i4", count=ne * 2, offset=off).reshape(ne, 2).astype(np.int64)
    off += ne * 8
    tris = np.frombuffer(d, dtype=">i4", count=nt * 3, offset=off).reshape(nt, 3).astype(np.int64)
    off += nt * 12
    tets = np.frombuffer(d, dtype=">i4", count=ntet * 4, offset=off).reshape(ntet, 4).astype(np.int64)
    off += ntet * 16
    coords = np.frombuffer(d, dtype=">f8", count=nv * 3, offset=off).reshape(nv, 3).astype(np.float64)
    # triangle vertices: edge 0 = (a, b); the third vertex is the one of edge 1 that is not on edge 0
    e0, e1 = edges[tris[:, 0]], edges[tris[:, 1]]
    third = np.where((e1[:, 0] != e0[:, 0]) & (e1[:, 0] != e0[:, 1]), e1[:, 0], e1[:, 1])
    tri_v = np.stack([e0[:, 0], e0[:, 1], third], axis=1)
    if dim == 2:
        if ntet:
            raise ValueError(f"{path}: 2D mesh with regions")
        return coords[:, :2].copy(), tri_v, 2
    if dim != 3:
        raise ValueError(f"{path}: dimension {dim} not supported")
    # tet vertices: the three of face 0 plus the vertex of face 1 that is not on face 0
    f0, f1 = tri_v[tets[:, 0]], tri_v[tets[:, 1]]
    on0 = (f1[:, :, None] == f0[:, None, :]).any(axis=2)            # [ntet, 3]: is vertex k of face 1 on face 0
    assert (on0.sum(axis=1) == 2).all(), "faces of a tet must share exactly one edge"
    fourth = f1[np.arange(ntet), np.argmin(on0, axis=1)]
    return coords.copy(), np.concatenate([f0, fourth[:, None]], axis=1), 3


def load_mesh(op, path, bc_of_face=None):
    """``Mesh`` (mesh.simplex_mesh) of a serial ``.smb`` file for the SBP operator ``op``."""
    from .mesh import simplex_mesh
    xyz, simp, dim = read_smb(path)
    if dim != op.dim:
        raise ValueError(f"{path} is a {dim}D mesh, the operator is {op.dim}D")
    v = xyz[simp]
    vol = np.linalg.det((v[:, 1:, :] - v[:, :1, :]).transpose(0, 2, 1))
    if not (np.abs(vol) > 0).all():
        raise ValueError(f"{path}: degenerate elements (some fixtures keep their coordinates outside the point block)")
    return simplex_mesh(op, xyz, simp, bc_of_face)
