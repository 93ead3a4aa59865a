"""ingestion of the reference's PUMI meshes (pdesolver.jl_b200/smb.py + mesh.simplex_mesh) and
the equivalence of the synthetic benchmark meshes with the reference's own benchmark meshes."""
import os

import numpy as np
import pytest

import oracle
import pdesolver_jl_b200 as pd
from pdesolver_jl_b200 import mesh as pmesh
from pdesolver_jl_b200 import smb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference/src/mesh_files"
C1 = {"Flux_name": "RoeFlux", "BC1_name": "isentropicVortexBC", "use_itermax": False}
C3 = {"Flux_name": "RoeFlux", "BC1_name": "ExpBC", "SRCname": "SRCExp", "use_itermax": False}


def fixture_mesh(op, name):
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    return pmesh.simplex_mesh(op, fx["vertex_coords"], fx["triangles" if op.dim == 2 else "tets"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
@pytest.mark.parametrize("name", ["squarevortex_small", "squarevortex_large", "square_benchmarksmall", "cube_benchmarksmall"])
def test_reader_reproduces_fixtures(name):
    xyz, simp, dim = smb.read_smb(os.path.join(REF, name + "0.smb"))
    fx = np.load(os.path.join(GOLD, name + ".npz"))
    assert np.array_equal(xyz, fx["vertex_coords"]) and np.array_equal(simp, fx["triangles" if dim == 2 else "tets"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_reader_header_counts_and_errors(tmp_path):
    # square_benchmarksmall 2601 vertices / 5000 triangles on [1,3]^2, cube_benchmarksmall 216 / 750 on [1.5,2.5]^3
    xy, tri, dim = smb.read_smb(os.path.join(REF, "square_benchmarksmall0.smb"))
    assert dim == 2 and xy.shape == (2601, 2) and tri.shape == (5000, 3) and abs(xy.min() - 1.0) < 1e-12 and abs(xy.max() - 3.0) < 1e-12
    xyz, tet, dim = smb.read_smb(os.path.join(REF, "cube_benchmarksmall0.smb"))
    assert dim == 3 and xyz.shape == (216, 3) and tet.shape == (750, 4) and abs(xyz.min() - 1.5) < 1e-12 and abs(xyz.max() - 2.5) < 1e-12
    with pytest.raises(ValueError):
        smb.load_mesh(pd.build_operator(2, 1), os.path.join(REF, "cube_benchmarksmall0.smb"))     # dimension mismatch
    with pytest.raises(ValueError):
        smb.load_mesh(pd.build_operator(2, 1), os.path.join(REF, "tri8l0.smb"))     # coordinates not in the point block
    bad = tmp_path / "bad.smb"
    bad.write_bytes(b"\x01" * 64)
    with pytest.raises(ValueError):
        smb.read_smb(str(bad))


@pytest.mark.parametrize("name,dim,p,volume", [("square_benchmarksmall", 2, 1, 4.0), ("cube_benchmarksmall", 3, 2, 1.0),
                                                ("cube_benchmarksmall", 3, 1, 1.0)])
def test_reference_mesh_is_conforming(name, dim, p, volume):
    op = pd.build_operator(dim, p)
    m = fixture_mesh(op, name)
    nf = dim + 1
    assert 2 * m.numInterfaces + m.numBoundaryFaces == nf * m.numEl       # every element face claimed exactly once
    assert abs((op.w[:, None] / m.jac).sum() - volume) < 1e-12            # sum of the mass matrix = domain volume
    # uniform flow: zero residual (test_dg.jl:115-128) -- normals, permutations and orientations are consistent
    opts = {"Flux_name": "RoeFlux", "BC1_name": "FreeStreamBC", "Ma": 0.4, "aoa": 10.0}
    P = oracle.Problem(m, op, opts)
    assert np.abs(P.eval_residual(P.exact_state("ICFreeStream"))).max() < 1e-12


@pytest.mark.parametrize("name,dim,p,n,ic,opts,h,kw", [
    ("square_benchmarksmall", 2, 1, 50, "ICIsentropicVortex", C1, 1e-3, {"diagonal": "\\"}),
    ("cube_benchmarksmall", 3, 2, 5, "ICExp", C3, 5e-5, {})])
def test_synthetic_benchmark_mesh_equals_reference_mesh(name, dim, p, n, ic, opts, h, kw):
    """perf/input_vals_2d_rk4.jl / perf/input_vals_3d_rk4.jl run on square_benchmarksmall / cube_benchmarksmall.  The
    structured meshes bench.py scales up are the same triangulations: the RK4 residual-norm history (a sum over all
    dofs, independent of PUMI's element numbering) agrees to round-off."""
    op = pd.build_operator(dim, p)
    hist = []
    for m in (fixture_mesh(op, name), pd.structured_mesh(op, n, **kw)):
        P = oracle.Problem(m, op, opts)
        _, _, norms = P.rk4(P.exact_state(ic), h, 4 * h)
        hist.append(norms)
    assert np.allclose(hist[0], hist[1], rtol=1e-9, atol=0)
