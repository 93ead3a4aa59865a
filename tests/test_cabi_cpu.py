"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/pdes_euler_b200.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

import pdesolver_jl_b200 as pd
from pdesolver_jl_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pdes_euler_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdes_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_cabi.LIB_PATH), "run __graft_entry__.build() first"
    L = C.CDLL(_cabi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(_cabi.EXPORTS) == names


def test_struct_layouts_match_header():
    assert C.sizeof(_cabi.PdesConfig) == 8 * 4 + 3 * 8 + 10 * 4 + 6 * 8
    assert pd.mesh.INTERFACE_DTYPE.itemsize == 12 and pd.mesh.BOUNDARY_DTYPE.itemsize == 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    op = pd.build_operator(2, 1)
    mesh = pd.structured_mesh(op, 2)
    with pytest.raises(pd.PDESolverError, match="no usable CUDA device"):
        pd.EulerData(mesh, op, {"Flux_name": "RoeFlux"})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pdesolver.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "liborc" not in txt, f


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("name", ["ICIsentropicVortex", "ICExp", "ICFreeStream"])
def test_host_ics_match_oracle(dim, name):
    import numpy as np
    import oracle
    from pdesolver_jl_b200 import ic
    op = pd.build_operator(dim, 1)
    mesh = pd.structured_mesh(op, 3)
    opts = {"Ma": 0.3, "aoa": 4.0}
    ref = oracle.Problem(mesh, op, opts).exact_state(name)
    got = ic.ICDict[name](mesh.coords, pd.ParamType(opts))
    assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()


def test_structured_mesh_per_dimension_counts():
    import numpy as np
    op = pd.build_operator(3, 1)
    m = pd.structured_mesh(op, (4, 2, 3))
    assert m.numEl == 4 * 2 * 3 * 6
    parts = (2, 1, 1)
    ms = [pd.structured_mesh(op, (4, 2, 3), parts=parts, rank=r) for r in range(2)]
    assert sum(x.numEl for x in ms) == m.numEl
    assert ms[0].peer_face_counts == ms[1].peer_face_counts == [2 * 3 * 2]
    assert m.numBoundaryFaces == 2 * 2 * (4 * 2 + 4 * 3 + 2 * 3)
