#!/usr/bin/env python
"""Converts the PUMI .smb meshes of the reference's convergence test (test/euler/convergence/p1/conservative_dg:
m1 = squarevortex_small, m2 = squarevortex_large) into small .npz fixtures (vertex coordinates + triangle vertex ids).

.smb layout (decoded in SURVEY.md §4): 48-byte big-endian header (magic, version, dim, nparts, entity counts), then the
downward adjacency (edge -> 2 vertices, triangle -> 3 edges, 4-byte ids), then numVert x 3 doubles.  Only this container
reads /root/reference; the tests read the .npz.

    python tests/golden/import_smb.py
"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/mesh_files"


def read_smb(path):
    d = open(path, "rb").read()
    h = struct.unpack(">12I", d[:48])
    assert h[1] == 4 and h[2] == 2, "expected a version-4 2D .smb"
    nv, ne, nt = h[4], h[5], h[6]
    off = 48
    edges = np.frombuffer(d, dtype=">i4", count=ne * 2, offset=off).reshape(ne, 2)
    off += ne * 8
    tris = np.frombuffer(d, dtype=">i4", count=nt * 3, offset=off).reshape(nt, 3)
    off += nt * 12
    coords = np.frombuffer(d, dtype=">f8", count=nv * 3, offset=off).reshape(nv, 3)
    tv = np.zeros((nt, 3), dtype=np.int64)
    for t in range(nt):
        e0, e1 = edges[tris[t, 0]], edges[tris[t, 1]]
        # edge 0 = (a, b); the third vertex is the one of edge 1 that is not in edge 0
        c = [v for v in e1 if v not in e0][0]
        tv[t] = (e0[0], e0[1], c)
    return coords[:, :2].astype(np.float64), tv


if __name__ == "__main__":
    for name in ("squarevortex_small", "squarevortex_large"):
        xy, tv = read_smb(os.path.join(SRC, name + "0.smb"))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), vertex_coords=xy, triangles=tv)
        print(name, xy.shape, tv.shape)
