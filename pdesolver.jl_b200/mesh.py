"""Host-side stand-in for PumiInterface.jl's DG mesh object (NOT vendored in
the reference) restricted to the fields the Euler hot path reads
(``docs/src/interfaces.md:191-596``):

``coords, dxidx, jac, interfaces, bndryfaces, bndry_offsets, nrm_face,
nrm_bndry, coords_bndry`` and the partition bookkeeping ``peer_parts,
peer_face_counts, bndries_local, shared_interfaces, nrm_sharedface``.

Synthetic structured meshes: ``n x n`` squares on [1,3]^2 cut
along the same diagonal into 2 triangles, or ``n^3`` cubes on [1.5,2.5]^3 cut
into 6 Kuhn tetrahedra; optional block partition ``parts=(px,py,pz)`` that
yields each rank's local mesh with shared-face lists ordered identically on
both sides of every partition boundary.

Conventions follow the reference: ``dxidx[k,p,j,e] = (d xi_k/d x_p)/jac``,
``jac = det(d xi/d x)``; ``nrm = dxidx^T * sbpface.normal[:,face]``
(``src/Utils/Utils.jl:666-683``); ``Interface(elementL, elementR, faceL, faceR,
orient)``; ``Boundary(element, face)``.  Indices are 0-based here.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np

from .sbp import SBPOperator, TRI_FACE_VTX, TET_FACE_VTX, TRI_VTX, TET_VTX

# 12-byte / 8-byte records, same layout as ODLCommonTools' Interface/Boundary
INTERFACE_DTYPE = np.dtype([("elementL", "<u4"), ("elementR", "<u4"),
                            ("faceL", "u1"), ("faceR", "u1"), ("orient", "u1"),
                            ("pad", "u1")])
BOUNDARY_DTYPE = np.dtype([("element", "<u4"), ("face", "u1"), ("pad", "u1", (3,))])


def _F(a):
    return np.asfortranarray(a)


@dataclass
class Mesh:
    dim: int
    numEl: int
    numNodesPerElement: int
    numNodesPerFace: int
    numDofPerNode: int
    coords: np.ndarray          # [dim, nn, nE]
    dxidx: np.ndarray           # [dim, dim, nn, nE]
    jac: np.ndarray             # [nn, nE]
    interfaces: np.ndarray      # INTERFACE_DTYPE [nF]
    bndryfaces: np.ndarray      # BOUNDARY_DTYPE [nB]
    bndry_offsets: np.ndarray   # [numBC+1]
    nrm_face: np.ndarray        # [dim, nfn, nF]
    nrm_bndry: np.ndarray       # [dim, nfn, nB]
    coords_bndry: np.ndarray    # [dim, nfn, nB]
    sbpface: object = None
    coord_order: int = 1
    isDG: bool = True
    # partition data (reference: docs/src/interfaces.md:573-596)
    myrank: int = 0
    commsize: int = 1
    peer_parts: list = field(default_factory=list)
    bndries_local: list = field(default_factory=list)       # per peer BOUNDARY_DTYPE
    shared_interfaces: list = field(default_factory=list)   # per peer INTERFACE_DTYPE
    nrm_sharedface: list = field(default_factory=list)      # per peer [dim,nfn,nfaces]
    shared_element_offsets: list = field(default_factory=list)
    remote_global_elnum: list = field(default_factory=list)  # per peer: global numbers of its elements in this part's halo
    local_element_lists: list = field(default_factory=list)  # per peer: my elements it needs, in ITS remote-element order
    global_elnum: np.ndarray = None                         # local -> global element id
    elem_vtx_coords: np.ndarray = None                      # [nE, dim+1, dim]

    @property
    def numInterfaces(self):
        return len(self.interfaces)

    @property
    def numBoundaryFaces(self):
        return len(self.bndryfaces)

    @property
    def numBC(self):
        return len(self.bndry_offsets) - 1

    @property
    def npeers(self):
        return len(self.peer_parts)

    @property
    def peer_face_counts(self):
        return [len(b) for b in self.bndries_local]

    @property
    def numDof(self):
        return self.numDofPerNode * self.numNodesPerElement * self.numEl


def _kuhn_tets():
    """6 positively oriented tets of the unit cube as corner offsets."""
    tets = []
    for perm in itertools.permutations(range(3)):
        p = np.zeros(3, dtype=int)
        path = [p.copy()]
        for ax in perm:
            p = p.copy()
            p[ax] += 1
            path.append(p)
        path = np.array(path)
        A = (path[1:] - path[0]).T
        if np.linalg.det(A) < 0:
            path[[1, 2]] = path[[2, 1]]
        tets.append(path)
    return np.array(tets)      # [6, 4, 3]


def block_ranges(n, parts, rank):
    """Cube-index range [lo, hi) per dimension owned by ``rank`` (``n`` cells per
    dimension: an int or one count per dimension)."""
    dim = len(parts)
    n = _per_dim(n, dim)
    idx = []
    r = rank
    for d in range(dim):
        idx.append(r % parts[d])
        r //= parts[d]
    lo = [(n[d] * idx[d]) // parts[d] for d in range(dim)]
    hi = [(n[d] * (idx[d] + 1)) // parts[d] for d in range(dim)]
    return lo, hi, idx


def _per_dim(n, dim):
    return [int(n)] * dim if np.isscalar(n) else [int(v) for v in n]


def _assemble(op, dim, gvid, vcoord, el_owner, gel, nE, rank, nranks, side_of):
    """Connectivity, metrics and partition lists from simplices given by global vertex ids ``gvid[ne, dim+1]``
    and vertex coordinates ``vcoord[ne, dim+1, dim]`` (local elements first).  ``side_of(face_centroids)`` maps
    boundary faces to (bc index array, numBC)."""
    ne_ext = len(gel)
    # ---- face pairing -----------------------------------------------------
    fvtx = TRI_FACE_VTX if dim == 2 else TET_FACE_VTX
    nfaces = dim + 1
    fv = gvid[:, fvtx]                                          # [ne_ext, nfaces, dim]
    key = np.sort(fv, axis=2).reshape(-1, dim)
    el_of = np.repeat(np.arange(ne_ext), nfaces)
    lf_of = np.tile(np.arange(nfaces), ne_ext)
    srt = np.lexsort(tuple(key[:, d] for d in range(dim - 1, -1, -1)))
    ks = key[srt]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    first = np.nonzero(same)[0]                                 # pair (first, first+1)
    paired = np.zeros(len(ks), dtype=bool)
    paired[first] = True
    paired[first + 1] = True
    a, b = srt[first], srt[first + 1]
    ea, eb = el_of[a], el_of[b]
    swap = ea > eb
    a, b = np.where(swap, b, a), np.where(swap, a, b)
    eL, eR, fL, fR = el_of[a], el_of[b], lf_of[a], lf_of[b]
    keyL = ks[first]

    if dim == 2:
        orient = np.zeros(len(eL), dtype=np.int64)
    else:
        vl = fv[eL, fL]                                         # [np, 3]
        vr = fv[eR, fR]
        orient = np.full(len(eL), -1, dtype=np.int64)
        orient[(vr[:, 0] == vl[:, 0]) & (vr[:, 1] == vl[:, 2])] = 0
        orient[(vr[:, 1] == vl[:, 1]) & (vr[:, 0] == vl[:, 2])] = 1
        orient[(vr[:, 2] == vl[:, 2]) & (vr[:, 0] == vl[:, 1])] = 2
        assert orient.min() >= 0, "inconsistent face orientation"

    locL, locR = eL < nE, eR < nE
    both = locL & locR
    shared = locL ^ locR

    # ---- element metrics ---------------------------------------------------
    nn = op.numnodes
    nfn = op.face.numnodes
    vloc = vcoord[:nE]                                          # [nE, dim+1, dim]
    A = 0.5 * (vloc[:, 1:, :] - vloc[:, :1, :]).transpose(0, 2, 1)   # dx/dxi [nE, dim, dim]
    detA = np.linalg.det(A)
    assert detA.min() > 0
    Ainv = np.linalg.inv(A)                                     # dxi/dx
    jac_e = 1.0 / detA
    dxidx_e = Ainv * detA[:, None, None]                        # (dxi/dx)/jac
    dxidx = np.empty((dim, dim, nn, nE), order="F")
    dxidx[:] = dxidx_e.transpose(1, 2, 0)[:, :, None, :]
    jac = np.empty((nn, nE), order="F")
    jac[:] = jac_e[None, :]
    coords = _F(np.einsum("jv,evd->dje", op.bary, vloc))

    ref_n = np.asarray(op.face.normal)                          # [dim, nfaces]
    fb = op.face.facenodes_bary                                 # [nfn, dim]

    def face_normals(els, lfaces, dx_src):
        # nrm_p = sum_k dxidx[k,p] * n_ref[k]
        nr = np.einsum("ekp,ke->pe", dx_src[els], ref_n[:, lfaces])
        out = np.empty((dim, nfn, len(els)), order="F")
        out[:] = nr[:, None, :]
        return out

    # interior interfaces
    iface = np.zeros(int(both.sum()), dtype=INTERFACE_DTYPE)
    iface["elementL"], iface["elementR"] = eL[both], eR[both]
    iface["faceL"], iface["faceR"], iface["orient"] = fL[both], fR[both], orient[both]
    io = np.lexsort((iface["faceL"], iface["elementL"]))
    iface = iface[io]
    nrm_face = face_normals(iface["elementL"].astype(np.int64),
                            iface["faceL"].astype(np.int64), dxidx_e)

    # boundary faces: unpaired faces of local elements
    un = srt[~paired]
    un = un[el_of[un] < nE]
    be, bf = el_of[un], lf_of[un]
    bvc = vcoord[be][np.arange(len(be))[:, None], fvtx[bf]]     # [nB, dim, dim] face vertex coords
    cen = bvc.mean(axis=1)
    bc, numBC = side_of(cen)
    bo = np.lexsort((bf, be, bc))
    be, bf, bc, bvc = be[bo], bf[bo], bc[bo], bvc[bo]
    bndry = np.zeros(len(be), dtype=BOUNDARY_DTYPE)
    bndry["element"], bndry["face"] = be, bf
    bndry_offsets = np.searchsorted(bc, np.arange(numBC + 1)).astype(np.int64)
    nrm_bndry = face_normals(be, bf, dxidx_e)
    coords_bndry = _F(np.einsum("iv,bvd->dib", fb, bvc))

    mesh = Mesh(dim=dim, numEl=nE, numNodesPerElement=nn, numNodesPerFace=nfn,
                numDofPerNode=dim + 2, coords=coords, dxidx=dxidx, jac=jac,
                interfaces=iface, bndryfaces=bndry, bndry_offsets=bndry_offsets,
                nrm_face=nrm_face, nrm_bndry=nrm_bndry, coords_bndry=coords_bndry,
                sbpface=op.face, myrank=rank, commsize=nranks,
                global_elnum=gel[:nE].copy(), elem_vtx_coords=vloc)

    # ---- shared faces ------------------------------------------------------
    if shared.any():
        # local element is always elementL of a shared interface
        sl = np.where(locL[shared], eL[shared], eR[shared])
        sr = np.where(locL[shared], eR[shared], eL[shared])
        sfl = np.where(locL[shared], fL[shared], fR[shared])
        sfr = np.where(locL[shared], fR[shared], fL[shared])
        so = orient[shared]
        skey = keyL[shared]
        peer = el_owner[sr]
        for p in np.unique(peer):
            m = peer == p
            # identical ordering on both ranks: sort by global face key
            kk = skey[m]
            o2 = np.lexsort(tuple(kk[:, d] for d in range(dim - 1, -1, -1)))
            bl = np.zeros(int(m.sum()), dtype=BOUNDARY_DTYPE)
            bl["element"], bl["face"] = sl[m][o2], sfl[m][o2]
            si = np.zeros(int(m.sum()), dtype=INTERFACE_DTYPE)
            si["elementL"], si["elementR"] = sl[m][o2], sr[m][o2]
            si["faceL"], si["faceR"], si["orient"] = sfl[m][o2], sfr[m][o2], so[m][o2]
            mesh.peer_parts.append(int(p))
            mesh.bndries_local.append(bl)
            mesh.shared_interfaces.append(si)
            mesh.nrm_sharedface.append(face_normals(sl[m][o2], sfl[m][o2], dxidx_e))
            mesh.shared_element_offsets.append(int(np.nonzero(el_owner == p)[0].min()))
            # global numbers of the peer's elements in this part's halo, in remote-element order: what the peer's
            # local_element_lists entry for this part must contain (negotiated once at mesh load, as PUMI does)
            mesh.remote_global_elnum.append(gel[el_owner == p].copy())
    return mesh


def structured_mesh(op: SBPOperator, n, parts=None, rank: int = 0,
                    domain=None, bc_sides=None, shuffle_seed=None, diagonal="/") -> Mesh:
    """Structured simplex mesh of ``n^dim`` cells; ``n`` may also be
    one cell count per dimension (the cell size then stays ``(domain[1]-domain[0])/n[0]``
    in every direction, i.e. the box grows: used for weak-scaling runs).

    parts: block partition (px,py[,pz]); ``rank`` selects the local block.
    bc_sides: optional list, one BC index per geometric side
    (xmin,xmax,ymin,ymax[,zmin,zmax]); default: all sides in BC 0.
    diagonal (2D): "/" cuts every square from (x0,y0) to (x1,y1); "\\" from (x1,y0) to (x0,y1), which is the
    triangulation of the reference's ``square_benchmarksmall`` (perf/input_vals_2d_rk4.jl; verified against the
    .smb fixture in tests/test_smb.py).  The 3D Kuhn split already equals ``cube_benchmarksmall``.
    shuffle_seed: if given, apply a pseudo-random even permutation (keyed on the
    global element number, so it is partition independent) to every element's
    vertex list; this exercises every faceL/faceR/orient combination, which the
    plain Kuhn ordering does not.
    """
    dim = op.dim
    if domain is None:
        domain = (1.0, 3.0) if dim == 2 else (1.5, 2.5)
    if parts is None:
        parts = (1,) * dim
    nranks = int(np.prod(parts))
    nv = _per_dim(n, dim)
    n = nv[0]
    lo, hi, _ = block_ranges(nv, parts, rank)
    # extended block: one ghost layer wherever a neighbour exists
    elo = [max(lo[d] - 1, 0) for d in range(dim)]
    ehi = [min(hi[d] + 1, nv[d]) for d in range(dim)]
    ext = [ehi[d] - elo[d] for d in range(dim)]

    # cells of the extended block, x fastest
    grids = np.meshgrid(*[np.arange(elo[d], ehi[d]) for d in range(dim)], indexing="ij")
    cells = np.stack([g.ravel(order="F") for g in grids], axis=1)     # [nc, dim]

    def owner_of(c):
        r = np.zeros(len(c), dtype=np.int64)
        mult = 1
        for d in range(dim):
            # inverse of block_ranges: part index whose [lo,hi) contains c
            pidx = np.zeros(len(c), dtype=np.int64)
            for k in range(parts[d]):
                a, b = (nv[d] * k) // parts[d], (nv[d] * (k + 1)) // parts[d]
                pidx[(c[:, d] >= a) & (c[:, d] < b)] = k
            r += mult * pidx
            mult *= parts[d]
        return r

    cell_owner = owner_of(cells)
    # drop ghost cells that are only corner/edge neighbours of other ranks: keep
    # them, they are harmless (their faces with local elements do not exist).

    if dim == 2:
        # A = (v00, v10, v11), B = (v00, v11, v01): same diagonal everywhere
        simp = np.array([[[0, 0], [1, 0], [1, 1]], [[0, 0], [1, 1], [0, 1]]])
        if diagonal == "\\":
            simp = np.array([[[0, 0], [1, 0], [0, 1]], [[1, 0], [1, 1], [0, 1]]])
        elif diagonal != "/":
            raise ValueError("diagonal must be '/' or '\\'")
    else:
        simp = _kuhn_tets()
    ns = simp.shape[0]
    # global vertex index grid coordinates of every element's vertices
    vgrid = cells[:, None, None, :] + simp[None, :, :, :]      # [nc, ns, dim+1, dim]
    vgrid = vgrid.reshape(-1, dim + 1, dim)
    el_owner = np.repeat(cell_owner, ns)
    mult = np.cumprod([1] + [nv[d] + 1 for d in range(dim - 1)]).astype(np.int64)
    gvid = (vgrid * mult).sum(axis=2)                           # [ne_ext, dim+1]
    h = (domain[1] - domain[0]) / n
    vcoord = domain[0] + h * vgrid.astype(np.float64)           # [ne_ext, dim+1, dim]
    # global element number (cell lexicographic, x fastest, then simplex)
    gcell = (cells * np.cumprod([1] + nv[:-1]).astype(np.int64)).sum(axis=1)
    gel = (gcell[:, None] * ns + np.arange(ns)[None, :]).ravel()

    if shuffle_seed is not None:
        if dim == 2:
            evenp = np.array([[0, 1, 2], [1, 2, 0], [2, 0, 1]])
        else:
            evenp = np.array([p for p in itertools.permutations(range(4))
                              if np.linalg.det(np.eye(4)[list(p)]) > 0])
        pick = ((gel * 2654435761 + shuffle_seed * 40503) >> 7) % len(evenp)
        sel = evenp[pick]                                        # [ne_ext, dim+1]
        gvid = np.take_along_axis(gvid, sel, axis=1)
        vcoord = np.take_along_axis(vcoord, sel[:, :, None], axis=1)
    # order: local elements first (in global order), then ghosts grouped by owner
    is_local = el_owner == rank
    order = np.lexsort((gel, np.where(is_local, -1, el_owner)))
    gvid, vcoord, el_owner, gel, is_local = (gvid[order], vcoord[order], el_owner[order],
                                             gel[order], is_local[order])
    nE = int(is_local.sum())
    ne_ext = len(gel)

    def side_of(cen):
        side = np.full(len(cen), -1, dtype=np.int64)
        for d in range(dim):
            side[np.abs(cen[:, d] - domain[0]) < 1e-9 * h + 1e-12] = 2 * d
            side[np.abs(cen[:, d] - (domain[0] + h * nv[d])) < 1e-9 * h + 1e-12] = 2 * d + 1
        assert side.min() >= 0, "unpaired face not on the domain boundary"
        sides = [0] * (2 * dim) if bc_sides is None else bc_sides
        return np.asarray(sides)[side], int(max(sides)) + 1

    mesh = _assemble(op, dim, gvid, vcoord, el_owner, gel, nE, rank, nranks, side_of)
    # mesh.local_element_lists (PumiInterface; read by getSendDataElement, Utils/parallel.jl:276-293): the peer's ghost
    # layer is every cell of its block extended by one, so it needs my elements in those cells, and it numbers them by
    # global element number (the order of its remote_global_elnum entry for this part)
    gcell_of = mesh.global_elnum // ns
    cw = np.cumprod([1] + nv[:-1]).astype(np.int64)
    cidx = np.stack([(gcell_of // cw[d]) % nv[d] for d in range(dim)], axis=1)
    for pr in mesh.peer_parts:
        plo, phi, _ = block_ranges(nv, parts, pr)
        inside = np.ones(len(cidx), dtype=bool)
        for d in range(dim):
            inside &= (cidx[:, d] >= max(plo[d] - 1, 0)) & (cidx[:, d] < min(phi[d] + 1, nv[d]))
        mesh.local_element_lists.append(np.nonzero(inside)[0].astype(np.int64))      # local elements are in global order
    return mesh


def two_element_mesh(op: SBPOperator) -> Mesh:
    """The 2-triangle square [-1,1]^2 of ``test/euler/test_lowlevel.jl:9-72``
    (mesh file tri2l): element 1 = (-1,-1),(1,1),(-1,1), element 2 =
    (-1,-1),(1,-1),(1,1); interface (1,2,faceL=1,faceR=3); dxidx and jac as
    asserted there."""
    assert op.dim == 2
    v = np.array([[[-1.0, -1], [1, 1], [-1, 1]], [[-1.0, -1], [1, -1], [1, 1]]])
    nn, nfn, dim = op.numnodes, op.face.numnodes, 2
    A = 0.5 * (v[:, 1:, :] - v[:, :1, :]).transpose(0, 2, 1)
    detA = np.linalg.det(A)
    dxe = np.linalg.inv(A) * detA[:, None, None]
    dxidx = np.empty((2, 2, nn, 2), order="F")
    dxidx[:] = dxe.transpose(1, 2, 0)[:, :, None, :]
    jac = np.empty((nn, 2), order="F")
    jac[:] = (1.0 / detA)[None, :]
    coords = _F(np.einsum("jv,evd->dje", op.bary, v))
    iface = np.zeros(1, dtype=INTERFACE_DTYPE)
    iface[0] = (0, 1, 0, 2, 0, 0)
    b = np.zeros(4, dtype=BOUNDARY_DTYPE)
    for k, (e, f) in enumerate([(0, 2), (1, 0), (0, 1), (1, 1)]):
        b[k]["element"], b[k]["face"] = e, f
    ref_n = np.asarray(op.face.normal)
    fb = op.face.facenodes_bary

    def nrm(els, fs):
        out = np.empty((2, nfn, len(els)), order="F")
        out[:] = np.einsum("ekp,ke->pe", dxe[els], ref_n[:, fs])[:, None, :]
        return out
    be, bf = b["element"].astype(int), b["face"].astype(int)
    bvc = v[be][np.arange(4)[:, None], TRI_FACE_VTX[bf]]
    return Mesh(dim=2, numEl=2, numNodesPerElement=nn, numNodesPerFace=nfn,
                numDofPerNode=4, coords=coords, dxidx=dxidx, jac=jac, interfaces=iface,
                bndryfaces=b, bndry_offsets=np.array([0, 4], dtype=np.int64),
                nrm_face=nrm(np.array([0]), np.array([0])), nrm_bndry=nrm(be, bf),
                coords_bndry=_F(np.einsum("iv,bvd->dib", fb, bvc)), sbpface=op.face,
                elem_vtx_coords=v)


def simplex_mesh(op: SBPOperator, vertex_coords, simplices, bc_of_face=None) -> Mesh:
    """Mesh object for an arbitrary conforming simplex mesh: ``vertex_coords[nV, dim]``, ``simplices[nE, dim+1]``
    (vertex ids).  Elements are re-oriented to positive volume.  ``bc_of_face(centroids) -> (bc index, numBC)``
    classifies boundary faces (default: one BC).  Used to run the reference's own mesh fixtures
    (tests/golden/import_smb.py converts PUMI .smb files)."""
    dim = op.dim
    simplices = np.array(simplices, dtype=np.int64)
    vc = np.asarray(vertex_coords, dtype=np.float64)[:, :dim]
    vcoord = vc[simplices]                                       # [nE, dim+1, dim]
    A = (vcoord[:, 1:, :] - vcoord[:, :1, :]).transpose(0, 2, 1)
    neg = np.linalg.det(A) < 0
    simplices[neg] = simplices[neg][:, [1, 0] + list(range(2, dim + 1))]
    vcoord = vc[simplices]
    nE = len(simplices)
    if bc_of_face is None:
        def bc_of_face(cen):
            return np.zeros(len(cen), dtype=np.int64), 1
    return _assemble(op, dim, simplices, vcoord, np.zeros(nE, dtype=np.int64), np.arange(nE), nE, 0, 1, bc_of_face)
