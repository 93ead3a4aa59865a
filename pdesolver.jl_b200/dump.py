"""Flat dump of one ``evalResidual`` call of the Julia reference (SURVEY.md §8(c) "Consequence").

The reference cannot run in the build image (Julia 0.6 + three un-vendored packages).  Wherever it does run,
``tools/dump_pdesolver.jl`` writes everything the hot path reads -- the SummationByParts operator
(``Q, w, interp, perm, nbrperm, wface``), the PumiInterface mesh arrays (``dxidx, jac, coords, nrm_face, nrm_bndry,
coords_bndry, interfaces, bndryfaces, bndry_offsets``), the options that select the functors, the state ``q`` and the
residual ``res`` Julia computed -- into ONE file in the format below.  ``load_dump`` turns such a file back into the
``(mesh, sbp, opts, q, res)`` objects the oracle and the CUDA path accept: that pins the operator VALUES and gives true
parity against Julia.  ``save_dump`` writes the same format from the Python side (round-trip tests).

File format (little endian, column-major arrays, indices 1-based as Julia holds them)::

    magic    8 bytes  "PDSDUMP1"
    nrecords int32
    record:  int32 name_len | name (ascii) | int32 dtype (1 f64, 2 i64, 3 utf-8 bytes) | int32 ndims | int64 dims[ndims] | data

Records: ``dim numDofPerNode`` (i64 scalars); operator ``Q[nn,nn,dim] w[nn] interp[ss,nfn] perm[ss|nfn,numfaces]
nbrperm[nfn,norient] wface[nfn] sparse_face``; mesh ``coords dxidx jac nrm_face nrm_bndry coords_bndry
interfaces[5,nF] (elementL, elementR, faceL, faceR, orient) bndryfaces[2,nB] (element, face) bndry_offsets[numBC+1]``;
``opts`` (utf-8, ``key=value`` lines); ``q[nd,nn,nE] res[nd,nn,nE]``; optional ``t``.
"""
from __future__ import annotations

import struct

import numpy as np

from .mesh import BOUNDARY_DTYPE, INTERFACE_DTYPE, Mesh
from .sbp import SBPFace, SBPOperator

MAGIC = b"PDSDUMP1"
_F64, _I64, _TXT = 1, 2, 3


def _write_record(fh, name, arr):
    nb = name.encode("ascii")
    fh.write(struct.pack("<i", len(nb)))
    fh.write(nb)
    if isinstance(arr, (bytes, str)):
        raw = arr.encode("utf-8") if isinstance(arr, str) else arr
        fh.write(struct.pack("<iiq", _TXT, 1, len(raw)))
        fh.write(raw)
        return
    a = np.asarray(arr)
    code = _I64 if np.issubdtype(a.dtype, np.integer) else _F64
    a = np.asfortranarray(a, dtype=np.int64 if code == _I64 else np.float64)
    fh.write(struct.pack("<ii", code, a.ndim))
    fh.write(struct.pack(f"<{a.ndim}q", *a.shape))
    fh.write(a.tobytes(order="F"))


def read_records(path):
    """The raw records of a dump file: name -> ndarray (Fortran order) or str."""
    out = {}
    with open(path, "rb") as fh:
        if fh.read(8) != MAGIC:
            raise ValueError(f"{path}: not a PDSDUMP1 file")
        (nrec,) = struct.unpack("<i", fh.read(4))
        for _ in range(nrec):
            (nl,) = struct.unpack("<i", fh.read(4))
            name = fh.read(nl).decode("ascii")
            code, nd = struct.unpack("<ii", fh.read(8))
            dims = struct.unpack(f"<{nd}q", fh.read(8 * nd)) if nd else ()
            n = int(np.prod(dims)) if nd else 1
            if code == _TXT:
                out[name] = fh.read(n).decode("utf-8")
            else:
                dt = np.dtype("<f8") if code == _F64 else np.dtype("<i8")
                out[name] = np.frombuffer(fh.read(8 * n), dtype=dt).reshape(dims, order="F").copy(order="F")
    return out


def _parse_opts(text):
    opts = {}
    for line in text.splitlines():
        if "=" not in line:
            continue
        k, v = line.split("=", 1)
        k, v = k.strip(), v.strip()
        if v in ("true", "false", "True", "False"):
            opts[k] = v.lower() == "true"
        else:
            try:
                opts[k] = int(v)
            except ValueError:
                try:
                    opts[k] = float(v)
                except ValueError:
                    opts[k] = v
    return opts


def load_dump(path):
    """-> (mesh, sbp, opts, q, res): the objects ``oracle.Problem(mesh, sbp, opts)`` / ``EulerData(mesh, sbp, opts)`` take
    (0-based indices inside, like every other mesh of this package)."""
    r = read_records(path)
    dim = int(r["dim"].ravel()[0])
    Q, w = r["Q"], r["w"].ravel()
    nn = Q.shape[0]
    sparse = bool(int(r["sparse_face"].ravel()[0]))
    interp, wface = r["interp"], r["wface"].ravel()
    perm = (r["perm"] - 1).astype(np.int64)
    nbr = (r["nbrperm"] - 1).astype(np.int64)
    nfn = wface.size
    if nbr.ndim == 1:
        nbr = nbr.reshape(nfn, -1, order="F")
    face = SBPFace(numnodes=nfn, stencilsize=interp.shape[0], interp=np.asfortranarray(interp), perm=np.asfortranarray(perm),
                   nbrperm=np.asfortranarray(nbr), wface=wface.copy(), normal=None, sparse=sparse)
    sbp = SBPOperator(dim=dim, degree=int(r["degree"].ravel()[0]) if "degree" in r else -1,
                      kind="diage" if sparse else "omega", numnodes=nn, Q=np.asfortranarray(Q), w=w.copy(),
                      bary=None, xref=None, face=face)
    itf = r["interfaces"].astype(np.int64).reshape(5, -1, order="F")
    ifaces = np.zeros(itf.shape[1], dtype=INTERFACE_DTYPE)
    for row, name in enumerate(("elementL", "elementR", "faceL", "faceR", "orient")):
        ifaces[name] = itf[row] - 1
    bf = r["bndryfaces"].astype(np.int64).reshape(2, -1, order="F")
    bfaces = np.zeros(bf.shape[1], dtype=BOUNDARY_DTYPE)
    bfaces["element"], bfaces["face"] = bf[0] - 1, bf[1] - 1
    q = r["q"]
    nd = q.shape[0]
    mesh = Mesh(dim=dim, numEl=q.shape[2], numNodesPerElement=nn, numNodesPerFace=nfn, numDofPerNode=nd,
                coords=r["coords"], dxidx=r["dxidx"], jac=r["jac"], interfaces=ifaces, bndryfaces=bfaces,
                bndry_offsets=(r["bndry_offsets"].ravel() - 1).astype(np.int64), nrm_face=r["nrm_face"],
                nrm_bndry=r["nrm_bndry"], coords_bndry=r["coords_bndry"], sbpface=face)
    opts = _parse_opts(r.get("opts", ""))
    return mesh, sbp, opts, np.asfortranarray(q), np.asfortranarray(r["res"]) if "res" in r else None


def save_dump(path, mesh, sbp, opts, q, res, t=0.0):
    """Writes the schema ``tools/dump_pdesolver.jl`` writes (indices back to 1-based)."""
    f = sbp.face
    itf = np.stack([mesh.interfaces[n].astype(np.int64) + 1 for n in ("elementL", "elementR", "faceL", "faceR", "orient")])
    bf = np.stack([mesh.bndryfaces["element"].astype(np.int64) + 1, mesh.bndryfaces["face"].astype(np.int64) + 1])
    recs = [("dim", np.array([mesh.dim])), ("numDofPerNode", np.array([mesh.numDofPerNode])),
            ("degree", np.array([sbp.degree])), ("Q", sbp.Q), ("w", sbp.w), ("interp", f.interp),
            ("perm", np.asarray(f.perm, dtype=np.int64) + 1), ("nbrperm", np.asarray(f.nbrperm, dtype=np.int64) + 1),
            ("wface", f.wface), ("sparse_face", np.array([int(f.sparse)])),
            ("coords", mesh.coords), ("dxidx", mesh.dxidx), ("jac", mesh.jac), ("nrm_face", mesh.nrm_face),
            ("nrm_bndry", mesh.nrm_bndry), ("coords_bndry", mesh.coords_bndry), ("interfaces", itf), ("bndryfaces", bf),
            ("bndry_offsets", np.asarray(mesh.bndry_offsets, dtype=np.int64) + 1),
            ("opts", "\n".join(f"{k}={str(v).lower() if isinstance(v, bool) else v}" for k, v in sorted(opts.items()))),
            ("q", q), ("res", res), ("t", np.array([float(t)]))]
    with open(path, "wb") as fh:
        fh.write(MAGIC)
        fh.write(struct.pack("<i", len(recs)))
        for name, arr in recs:
            _write_record(fh, name, arr)
