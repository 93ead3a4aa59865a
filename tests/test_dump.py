"""The Julia dump path: a PDSDUMP1 file written by ``tools/dump_pdesolver.jl`` (or, here, by the Python
writer of the same schema) is turned back into mesh / operator / options objects; the oracle -- and on a GPU the CUDA
path -- must reproduce the dumped residual.  ``--dump FILE`` (or PDES_DUMP=FILE) checks a file that came from Julia."""
import os

import numpy as np
import pytest

import oracle
import pdesolver_jl_b200 as pd
from pdesolver_jl_b200 import dump
from common import CASES, KIND, perturbed, rel_l2

DUMP_CASES = ["3d_p2_alt_roe_src", "c1_2d_p1_roe", "3d_p1_es"]


def _write(tmp_path, case, n=3):
    dim, p, ic, opts = CASES[case]
    op = pd.build_operator(dim, p, KIND.get(case, "omega"))
    mesh = pd.structured_mesh(op, n, shuffle_seed=5)
    orc = oracle.Problem(mesh, op, opts)
    q0 = perturbed(orc.exact_state(ic), amp=1e-2)
    res = orc.eval_residual(q0)
    path = os.path.join(str(tmp_path), case + ".pds")
    dump.save_dump(path, mesh, op, opts, q0, res)
    return path, mesh, op, opts, q0, res


@pytest.mark.parametrize("case", DUMP_CASES)
def test_dump_round_trip_oracle(tmp_path, case):
    path, mesh, op, opts, q0, res = _write(tmp_path, case)
    m2, s2, o2, q2, r2 = dump.load_dump(path)
    assert np.array_equal(q2, q0) and np.array_equal(r2, res)
    assert np.array_equal(s2.Q, op.Q) and np.array_equal(s2.face.perm, op.face.perm)
    assert np.array_equal(m2.interfaces, mesh.interfaces) and np.array_equal(m2.bndryfaces, mesh.bndryfaces)
    for k, v in opts.items():
        assert o2[k] == v, (k, o2[k], v)
    # the oracle built from the dump alone reproduces the dumped residual
    assert rel_l2(oracle.Problem(m2, s2, o2).eval_residual(q2), r2) < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("case", DUMP_CASES)
def test_dump_round_trip_gpu(tmp_path, case):
    """A non-template operator size among the cases: the dump loader + the size-generic kernels accept whatever operator the
    Julia side hands over."""
    path = _write(tmp_path, case)[0]
    m2, s2, o2, q2, r2 = dump.load_dump(path)
    eqn = pd.EulerData(m2, s2, o2)
    eqn.q[...] = q2
    pd.evalResidual(m2, s2, eqn, o2)
    assert rel_l2(eqn.res, r2) < 1e-12


def _julia_dump():
    return os.environ.get("PDES_DUMP")


@pytest.mark.skipif(not _julia_dump(), reason="set PDES_DUMP=<file written by tools/dump_pdesolver.jl> to check a Julia dump")
def test_julia_dump_oracle():
    m, s, o, q, r = dump.load_dump(_julia_dump())
    assert rel_l2(oracle.Problem(m, s, o).eval_residual(q), r) < 1e-12


@pytest.mark.gpu
@pytest.mark.skipif(not _julia_dump(), reason="set PDES_DUMP=<file written by tools/dump_pdesolver.jl> to check a Julia dump")
def test_julia_dump_gpu():
    m, s, o, q, r = dump.load_dump(_julia_dump())
    eqn = pd.EulerData(m, s, o)
    eqn.q[...] = q
    pd.evalResidual(m, s, eqn, o)
    assert rel_l2(eqn.res, r) < 1e-12


def test_julia_dump_script_matches_the_schema():
    """tools/dump_pdesolver.jl cannot run here (no Julia): at least its record count and record names must agree with what the
    Python writer of the same schema emits, so that load_dump finds every record it needs."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    jl = open(os.path.join(root, "tools", "dump_pdesolver.jl")).read()
    body = jl[jl.index('write(io, "PDSDUMP1")'):]
    declared = int(re.search(r'write\(io, "PDSDUMP1"\); write\(io, Int32\((\d+)\)\)', body).group(1))
    names = re.findall(r'wrec\(io, "([A-Za-z_]+)"', body)
    assert declared == len(names), (declared, names)
    op = pd.build_operator(2, 1)
    mesh = pd.structured_mesh(op, 2)
    opts = {"Flux_name": "RoeFlux", "BC1_name": "isentropicVortexBC"}
    q = np.zeros((4, 3, mesh.numEl), order="F")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "x.pds")
        dump.save_dump(f, mesh, op, opts, q, q)
        assert sorted(dump.read_records(f)) == sorted(names)
