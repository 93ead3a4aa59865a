"""Worker for the multi-process tests (launched by torchrun or mp.spawn).

mode "gloo-oracle": CPU, gloo backend -- each rank owns one block of the partitioned mesh, packs its shared-face
    states with the oracle's getSendDataFace, exchanges them with dist.send/recv (the MPI Isend/Irecv of
    Utils/parallel.jl:82-141) and evaluates the oracle residual; the result must equal the serial one.
mode "nccl-b200": one GPU per rank -- the same comparison for the CUDA path with the NCCL halo exchange inside
    libpdes_euler_b200.so, for evalResidual and for an RK4 trajectory (norms carry the reference's sqrt(P) quirk).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PARTS = {2: {2: (2, 1), 3: (2, 1, 1)}, 4: {2: (2, 2), 3: (2, 2, 1)}, 8: {2: (4, 2), 3: (2, 2, 2)}}


def serial_and_local(case, n, rank, world, seed=4):
    import oracle
    import pdesolver_jl_b200 as pd
    from common import CASES, KIND, perturbed
    dim, p, ic, opts = CASES[case]
    op = pd.build_operator(dim, p, KIND.get(case, "omega"))
    parts = PARTS[world][dim]
    serial = pd.structured_mesh(op, n, shuffle_seed=seed)
    local = pd.structured_mesh(op, n, parts=parts, rank=rank, shuffle_seed=seed)
    orc_s = oracle.Problem(serial, op, opts)
    q_s = perturbed(orc_s.exact_state(ic), amp=1e-2 if case in KIND else 1e-3)
    pos = {int(g): i for i, g in enumerate(serial.global_elnum)}
    idx = np.array([pos[int(g)] for g in local.global_elnum])
    return pd, op, dict(opts), serial, local, orc_s, q_s, idx


def run_gloo_oracle(rank, world, case="c3_3d_p2_roe_src", n=4):
    import torch
    import torch.distributed as dist
    import oracle
    from common import rel_l2
    pd, op, opts, serial, local, orc_s, q_s, idx = serial_and_local(case, n, rank, world)
    orc = oracle.Problem(local, op, opts)
    q = np.asfortranarray(q_s[:, :, idx])
    orc.start_exchange(q)
    reqs, bufs = [], []
    for pi, pr in enumerate(local.peer_parts):
        send = torch.from_numpy(np.ascontiguousarray(orc.q_send[pi].ravel(order="F")))
        recv = torch.empty_like(send)
        bufs.append((pi, recv))
        reqs.append(dist.isend(send, dst=pr))
        reqs.append(dist.irecv(recv, src=pr))
    for r in reqs:
        r.wait()
    for pi, recv in bufs:
        orc.q_recv[pi][...] = recv.numpy().reshape(orc.q_recv[pi].shape, order="F")
    res = orc.eval_residual(q)
    err = rel_l2(res, orc_s.eval_residual(q_s)[:, :, idx])
    assert err < 1e-13, f"rank {rank}: partitioned oracle != serial oracle ({err:.2e})"
    # global norm through the process group == serial norm (calcNorm + Allreduce, Utils.jl:427-449)
    M = 1.0 / orc.mass_matrix_inverse()
    loc = torch.tensor([float(np.sum(res * M * res))], dtype=torch.float64)
    dist.all_reduce(loc)
    Ms = 1.0 / orc_s.mass_matrix_inverse()
    rs = orc_s.eval_residual(q_s)
    assert abs(loc.item() - float(np.sum(rs * Ms * rs))) < 1e-12 * loc.item()


def run_gloo_oracle_es2(rank, world, dim=3, p=1, n=3):
    """face_integral_type = 2 on a partitioned mesh (element-data halo): each rank tells its peers which
    of their elements it needs (the negotiation PUMI does at mesh load), every rank sends those elements whole
    (getSendDataElement, Utils/parallel.jl:276-293) and evaluates calcSharedFaceElementIntegrals_element_inner
    (flux.jl:442-496); the result must equal the serial one."""
    import torch
    import torch.distributed as dist
    import oracle
    import pdesolver_jl_b200 as pd
    from common import perturbed, rel_l2
    op = pd.build_operator(dim, p)
    parts = PARTS[world][dim]
    serial = pd.structured_mesh(op, n, shuffle_seed=4)
    local = pd.structured_mesh(op, n, parts=parts, rank=rank, shuffle_seed=4)
    ic, bc = ("ICIsentropicVortex", "isentropicVortexBC") if dim == 2 else ("ICExp", "ExpBC")
    opts = {"Flux_name": "IRFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2, "face_integral_type": 2,
            "FaceElementIntegral_name": "ESLFFaceIntegral", "BC1_name": bc}
    orc_s = oracle.Problem(serial, op, opts)
    q_s = perturbed(orc_s.exact_state(ic), amp=1e-2)
    pos = {int(g): i for i, g in enumerate(serial.global_elnum)}
    idx = np.array([pos[int(g)] for g in local.global_elnum])
    q = np.asfortranarray(q_s[:, :, idx])
    orc = oracle.Problem(local, op, opts)
    mine = {int(g): i for i, g in enumerate(local.global_elnum)}
    # 1. negotiation: the global numbers of the peer's elements in my halo, in my remote-element order
    want, reqs = {}, []
    for pi, pr in enumerate(local.peer_parts):
        ask = torch.from_numpy(np.ascontiguousarray(local.remote_global_elnum[pi], dtype=np.int64))
        cnt = torch.tensor([len(ask)], dtype=torch.int64)
        other = torch.zeros(1, dtype=torch.int64)
        r1, r2 = dist.isend(cnt, dst=pr), dist.irecv(other, src=pr)
        r1.wait(); r2.wait()
        want[pi] = torch.zeros(int(other.item()), dtype=torch.int64)
        reqs += [dist.isend(ask, dst=pr), dist.irecv(want[pi], src=pr)]
    for r in reqs:
        r.wait()
    local_element_lists = {pi: [mine[int(g)] for g in w.numpy()] for pi, w in want.items()}
    # 2. exchange of whole elements
    reqs, bufs = [], []
    for pi, pr in enumerate(local.peer_parts):
        send = torch.from_numpy(np.ascontiguousarray(orc.get_send_data_element(q, local_element_lists[pi]).ravel(order="F")))
        nrem = len(local.remote_global_elnum[pi])
        recv = torch.empty(q.shape[0] * q.shape[1] * nrem, dtype=torch.float64)
        bufs.append((pi, recv, nrem))
        reqs += [dist.isend(send, dst=pr), dist.irecv(recv, src=pr)]
    for r in reqs:
        r.wait()
    for pi, recv, nrem in bufs:
        orc.set_recv_elements(pi, recv.numpy().reshape((q.shape[0], q.shape[1], nrem), order="F"))
    res = orc.eval_residual(q)
    err = rel_l2(res, orc_s.eval_residual(q_s)[:, :, idx])
    assert err < 1e-13, f"rank {rank}: partitioned type-2 oracle != serial oracle ({err:.2e})"


def run_nccl_b200(rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from common import rel_l2
    torch.cuda.set_device(local_rank)
    for case, n, h in [("c3_3d_p2_roe_src", 4, 5e-5), ("c1_2d_p1_roe", 8, 1e-3), ("3d_p1_roe_src", 4, 5e-5),
                       ("c2_2d_p2_es", 6, 1e-3)]:
        pd, op, opts, serial, local, orc_s, q_s, idx = serial_and_local(case, n, rank, world)
        ids = [pd.EulerData.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eqn = pd.EulerData(local, op, opts, device=local_rank, comm=(ids[0], rank, world))
        eqn.q[...] = q_s[:, :, idx]
        pd.evalResidual(local, op, eqn, opts)
        err = rel_l2(eqn.res, orc_s.eval_residual(q_s)[:, :, idx])
        assert err < 1e-12, f"rank {rank} {case}: residual parity {err:.2e}"
        nsteps = 10
        opts["use_itermax"] = False
        eqn.q[...] = q_s[:, :, idx]
        t = pd.rk4(pd.evalResidual, h, nsteps * h, local, op, eqn, opts)
        t_ref, q_ref, norms_ref = orc_s.rk4(q_s, h, nsteps * h)
        errq = rel_l2(eqn.q, q_ref[:, :, idx])
        assert t == t_ref and errq < 1e-10, f"rank {rank} {case}: rk4 parity {errq:.2e}"
        # the parallel norm is reduced twice -> sqrt(P) * serial norm
        assert np.allclose(eqn.convergence, np.sqrt(world) * norms_ref, rtol=1e-11, atol=0), \
            f"rank {rank} {case}: norms"
        # res_tol exit on a partitioned mesh (rk4.jl:258-267): every rank must stop at the same step head.  The norm the
        # reference tests in parallel is the doubly reduced one (sqrt(P) x the serial norm, rk4.jl:451-453), so the serial
        # oracle runs with res_tol / sqrt(P); lserk54 through the same halo.
        if case == "c1_2d_p1_roe":
            tol = float(np.sqrt(world) * 0.5 * (norms_ref[4] + norms_ref[5]))
            eqn.q[...] = q_s[:, :, idx]
            t = pd.rk4(pd.evalResidual, h, 40 * h, local, op, eqn, opts, res_tol=tol)
            t_ref, q_ref, n_ref = orc_s.rk4(q_s, h, 40 * h, res_tol=tol / np.sqrt(world))
            assert t == t_ref and len(eqn.convergence) == len(n_ref) < 40, (t, t_ref, len(eqn.convergence), len(n_ref))
            errq = rel_l2(eqn.q, q_ref[:, :, idx])
            assert errq < 1e-10, f"rank {rank} {case}: res_tol exit {errq:.2e}"
            eqn.q[...] = q_s[:, :, idx]
            t = pd.lserk54(pd.evalResidual, h, 8 * h, local, op, eqn, opts)
            t_ref, q_ref, n_ref = orc_s.lserk54(q_s, h, 8 * h)
            errq = rel_l2(eqn.q, q_ref[:, :, idx])
            assert t == t_ref and errq < 1e-10, f"rank {rank} {case}: lserk54 {errq:.2e}"
            assert np.allclose(eqn.convergence, np.sqrt(world) * n_ref, rtol=1e-11, atol=0)
        if case not in ("c2_2d_p2_es",):
            # J*v and the Newton-Krylov linear solve on the partitioned mesh (newton_setup.jl:632-662 is flux- and
            # partition-agnostic): the shared-face states AND directions are exchanged, the Krylov inner products are
            # all-reduced.  Reference: the same product / solve by a serial context on this rank's GPU.
            es = pd.EulerData(serial, op, opts, device=local_rank)
            rng = np.random.RandomState(7)
            v_s = np.asfortranarray(rng.standard_normal(q_s.shape) * np.abs(q_s))
            es.q[...] = q_s
            eqn.q[...] = q_s[:, :, idx]
            Jv_s = pd.evaldRdqProduct(serial, op, es, opts, v_s)
            Jv_p = pd.evaldRdqProduct(local, op, eqn, opts, np.asfortranarray(v_s[:, :, idx]))
            errj = rel_l2(Jv_p, Jv_s[:, :, idx])
            assert errj < 1e-12, f"rank {rank} {case}: partitioned J*v {errj:.2e}"
            kopts = dict(opts, krylov_reltol=1e-10, krylov_itermax=300, krylov_restart=100)
            b_s = np.asfortranarray(rng.standard_normal(q_s.shape))
            x_s = pd.linearSolve(serial, op, es, kopts, b_s)
            x_p = pd.linearSolve(local, op, eqn, kopts, np.asfortranarray(b_s[:, :, idx]))
            ks, kp = es.krylov_info, eqn.krylov_info
            assert kp["reason"] == ks["reason"] and abs(kp["iterations"] - ks["iterations"]) <= 2, (kp, ks)
            if ks["reason"] > 0:
                errx = rel_l2(x_p, x_s[:, :, idx])
                assert errx < 1e-6, f"rank {rank} {case}: partitioned GMRES {errx:.2e}"
            else:
                assert abs(kp["rnorm"] - ks["rnorm"]) < 1e-6 * ks["rnorm"], (kp, ks)
            es.close()
        eqn.close()
        dist.barrier()
    # type-2 entropy-stable face integrals through the element-data halo (ncclSend/ncclRecv of whole elements)
    import oracle
    for dim, p, n, fei in [(3, 1, 3, "ESLFFaceIntegral"), (2, 2, 4, "ESLW2FaceIntegral")]:
        import pdesolver_jl_b200 as pd
        from common import perturbed
        op = pd.build_operator(dim, p)
        parts = PARTS[world][dim]
        serial = pd.structured_mesh(op, n, shuffle_seed=4)
        local = pd.structured_mesh(op, n, parts=parts, rank=rank, shuffle_seed=4)
        ic, bc = ("ICIsentropicVortex", "isentropicVortexBC") if dim == 2 else ("ICExp", "ExpBC")
        opts = {"Flux_name": "IRFlux", "Volume_flux_name": "IRFlux", "volume_integral_type": 2, "face_integral_type": 2,
                "FaceElementIntegral_name": fei, "BC1_name": bc, "use_itermax": False}
        orc_s = oracle.Problem(serial, op, opts)
        q_s = perturbed(orc_s.exact_state(ic), amp=1e-2)
        pos = {int(g): i for i, g in enumerate(serial.global_elnum)}
        idx = np.array([pos[int(g)] for g in local.global_elnum])
        ids = [pd.EulerData.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        eqn = pd.EulerData(local, op, opts, device=local_rank, comm=(ids[0], rank, world))
        eqn.q[...] = q_s[:, :, idx]
        pd.evalResidual(local, op, eqn, opts)
        err = rel_l2(eqn.res, orc_s.eval_residual(q_s)[:, :, idx])
        assert err < 1e-12, f"rank {rank} type-2 {dim}D p{p}: residual parity {err:.2e}"
        h = 1e-3 if dim == 2 else 5e-5
        eqn.q[...] = q_s[:, :, idx]
        pd.rk4(pd.evalResidual, h, 5 * h, local, op, eqn, opts)
        _, q_ref, _ = orc_s.rk4(q_s, h, 5 * h)
        errq = rel_l2(eqn.q, q_ref[:, :, idx])
        assert errq < 1e-10, f"rank {rank} type-2 {dim}D p{p}: rk4 parity {errq:.2e}"
        eqn.close()
        dist.barrier()
    if rank == 0:
        print(f"nccl-b200 ok on {world} ranks")


def main():
    import torch.distributed as dist
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if mode == "gloo-oracle":
        dist.init_process_group("gloo")
        run_gloo_oracle(rank, world)
    elif mode == "gloo-oracle-es2":
        dist.init_process_group("gloo")
        run_gloo_oracle_es2(rank, world, dim=3, p=1, n=3)
        run_gloo_oracle_es2(rank, world, dim=2, p=2, n=4)
    else:
        import torch
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        run_nccl_b200(rank, world, local_rank)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
