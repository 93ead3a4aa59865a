#!/usr/bin/env python
"""Converts PUMI .smb meshes of the reference into small .npz fixtures (vertex coordinates + simplex vertex ids) with the
package's reader (pdesolver.jl_b200/smb.py):

  squarevortex_small / squarevortex_large   test/euler/convergence/p1/conservative_dg (m1, m2)
  square_benchmarksmall                     perf/input_vals_2d_rk4.jl   (BASELINE.json configuration 1: 5000 triangles)
  cube_benchmarksmall                       perf/input_vals_3d_rk4.jl   (configuration 3's reference mesh: 750 tets)

Only this container reads /root/reference; the tests read the .npz.

    python tests/golden/import_smb.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SRC = "/root/reference/src/mesh_files"
NAMES = ("squarevortex_small", "squarevortex_large", "square_benchmarksmall", "cube_benchmarksmall")

if __name__ == "__main__":
    from pdesolver_jl_b200 import smb
    for name in NAMES:
        xyz, simp, dim = smb.read_smb(os.path.join(SRC, name + "0.smb"))
        key = "triangles" if dim == 2 else "tets"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), vertex_coords=xyz, **{key: simp})
        print(name, xyz.shape, simp.shape)
