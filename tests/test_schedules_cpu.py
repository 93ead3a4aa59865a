"""Host-side checks of two schedules the CUDA kernels rely on (no GPU needed):

* the round-robin pair schedule of k_element_split_r (es_kernels.cuh): nn/2 rounds of disjoint node pairs must cover every
  pair {i, m} of an element exactly once, for odd and even nn;
* the RK4 stage update without the running sum of the k's (residual_kernels.cuh, epilogue_tile scheme 2): the regrouped
  update must equal the reference's x + h/6 (k1 + 2 k2 + 2 k3 + k4) (NonlinearSolvers/rk4.jl:244-319) to rounding."""
import numpy as np
import pytest


@pytest.mark.parametrize("nn", [3, 4, 6, 11, 12])
def test_round_robin_rounds_cover_every_pair_once(nn):
    seen = {}
    for k in range(1, nn // 2 + 1):
        half = nn % 2 == 0 and k == nn // 2
        senders = set()
        for i in range(nn):
            if half and i >= nn // 2:
                continue
            m = (i + k) % nn
            pair = (min(i, m), max(i, m))
            seen[pair] = seen.get(pair, 0) + 1
            senders.add(i)
        # the receiver of round k is node i + k: it reads the slot of node (i' - k) mod nn, which must have evaluated
        for ip in range(nn):
            src = (ip - k) % nn
            assert (src in senders) == (not half or src < nn // 2)
    assert len(seen) == nn * (nn - 1) // 2 and set(seen.values()) == {1}


def test_rk4_without_running_sum_equals_reference_form():
    rng = np.random.RandomState(3)
    n = 400
    A = rng.standard_normal((n, n)) / np.sqrt(n)
    src = rng.standard_normal(n)            # the tabulated, time-independent source Minv * srcw

    def minv_R(q):                          # Minv * R(q) without the source, as the kernel's accumulator holds it
        return A @ q + 0.1 * np.sin(q)

    x = rng.standard_normal(n)
    h = 1e-2
    # reference form
    k1 = minv_R(x) + src
    k2 = minv_R(x + 0.5 * h * k1) + src
    k3 = minv_R(x + 0.5 * h * k2) + src
    k4 = minv_R(x + h * k3) + src
    ref = x + (h / 6.0) * (k1 + 2 * k2 + 2 * k3 + k4)
    # scheme 2: q2, q3, the one extra vector w', q4, and the final combination (stage 4 reads neither x nor src)
    q2 = x + 0.5 * h * (minv_R(x) + src)
    q3 = x + 0.5 * h * (minv_R(q2) + src)
    w = q2 + 2.0 * q3 - x + 0.5 * h * src
    q4 = x + h * (minv_R(q3) + src)
    new = (w + q4) * (1.0 / 3.0) + (h / 6.0) * minv_R(q4)
    assert np.linalg.norm(new - ref) / np.linalg.norm(ref) < 5e-16 * 10


@pytest.mark.parametrize("dim,parts", [(2, (2, 1)), (2, (2, 2)), (2, (4, 2)), (3, (2, 1, 1)), (3, (2, 2, 1)), (3, (2, 2, 2))])
def test_partition_bookkeeping_is_symmetric(dim, parts):
    """What the halo transports rely on (pdes_api.cu setup_p2p, ncclSend/ncclRecv pairs, the element-data halo): if rank r
    lists peer p then p lists r with the same number of shared faces, in the same (global) face order, and the halo element
    lists a rank asks for are elements its peer owns.  Checked for every partition bench.py uses (2 / 4 / 8 ranks)."""
    import pdesolver_jl_b200 as pd
    op = pd.build_operator(dim, 1)
    nranks = int(np.prod(parts))
    n = tuple(3 * p for p in parts)
    meshes = [pd.structured_mesh(op, n, parts=parts, rank=r, shuffle_seed=2) for r in range(nranks)]
    total = 0
    for r, m in enumerate(meshes):
        assert len(set(m.peer_parts)) == len(m.peer_parts) and r not in m.peer_parts
        for pi, p in enumerate(m.peer_parts):
            o = meshes[p]
            assert r in o.peer_parts, f"rank {p} does not list rank {r}"
            po = o.peer_parts.index(r)
            assert len(m.bndries_local[pi]) == len(o.bndries_local[po]) > 0
            # same faces in the same order: the global numbers of (local element, remote element) mirror each other
            mine_l = m.global_elnum[m.shared_interfaces[pi]["elementL"]]
            mine_r = m.remote_global_elnum[pi][m.shared_interfaces[pi]["elementR"] - m.shared_element_offsets[pi]]
            theirs_l = o.global_elnum[o.shared_interfaces[po]["elementL"]]
            theirs_r = o.remote_global_elnum[po][o.shared_interfaces[po]["elementR"] - o.shared_element_offsets[po]]
            assert np.array_equal(mine_l, theirs_r) and np.array_equal(mine_r, theirs_l)
            assert np.array_equal(m.shared_interfaces[pi]["faceL"], o.shared_interfaces[po]["faceR"])
            assert np.array_equal(m.shared_interfaces[pi]["orient"], o.shared_interfaces[po]["orient"])
            # shared normals are equal and opposite
            assert np.allclose(m.nrm_sharedface[pi][:, 0, :], -o.nrm_sharedface[po][:, 0, :], atol=1e-14)
            assert set(m.remote_global_elnum[pi].tolist()) <= set(o.global_elnum.tolist())
            total += len(m.bndries_local[pi])
        assert len(m.peer_parts) <= 32          # HaloRec capacity of the peer-to-peer set-up
    assert total > 0 and total % 2 == 0


def test_documented_switches_exist_in_the_sources():
    """README.md lists the run-time switches: every PDES_* name it documents must be read somewhere in the library or the
    Python host, so that the table cannot drift from the code."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    readme = open(os.path.join(root, "README.md")).read()
    table = readme[readme.index("## Run-time switches"):]
    names = set(re.findall(r"`(PDES_[A-Z0-9_]+)", table))
    assert len(names) >= 15
    src = ""
    for d in ("pdesolver.jl_b200/csrc", "pdesolver.jl_b200"):
        for f in os.listdir(os.path.join(root, d)):
            if f.endswith((".cu", ".cuh", ".py")) or f == "Makefile":
                src += open(os.path.join(root, d, f)).read()
    missing = sorted(n for n in names if n not in src)
    assert not missing, f"documented but not read anywhere: {missing}"


def test_local_element_lists_match_the_peers_halo():
    """mesh.local_element_lists (getSendDataElement, Utils/parallel.jl:276-293): what a part sends to a peer is exactly, and
    in the same order, what that peer holds as its remote elements of this part."""
    import numpy as np
    import pdesolver_jl_b200 as pd
    for dim, p, parts in [(2, 1, (2, 2)), (3, 1, (2, 2, 2)), (3, 2, (2, 1, 1)), (2, 2, (4, 2))]:
        op = pd.build_operator(dim, p)
        nr = int(np.prod(parts))
        ms = [pd.structured_mesh(op, 5, parts=parts, rank=r, shuffle_seed=3) for r in range(nr)]
        for r, m in enumerate(ms):
            for pi, pr in enumerate(m.peer_parts):
                po = ms[pr].peer_parts.index(r)
                assert np.array_equal(m.global_elnum[m.local_element_lists[pi]], ms[pr].remote_global_elnum[po])
