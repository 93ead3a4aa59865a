"""Host-side stand-in for SummationByParts.jl (NOT vendored in the reference).

The reference obtains its SBP operators from the external package
SummationByParts.jl (branch ``jcwork``; constructors called at
``src/solver/common.jl:288-379``: getTriSBPOmega0/getTetSBPOmega/getTriSBPDiagE,
TriFace/TetFace/getTriFaceForDiagE).  Operator *values* are therefore inputs to
the hot path, not part of it: the Julia host passes ``sbp.Q, sbp.w,
sbpface.interp, perm, nbrperm, wface`` through the C ABI.

For the synthetic benchmarks and the parity tests (no Julia here) this module
builds valid multi-dimensional SBP operators from first principles with the
same array shapes and index conventions the reference uses
(conventions confirmed inside the reference at
``src/jacobian/jacobian.jl:1015-1125``, ``src/solver/euler/faceElementIntegrals.jl:81-103``
and ``test/euler/test_curvilinear.jl:22-60``):

* reference triangle (-1,-1),(1,-1),(-1,1); reference tet (-1,-1,-1),(1,-1,-1),
  (-1,1,-1),(-1,-1,1);
* ``Q[i,j,d]`` with ``Q_d + Q_d^T = E_d``, ``w`` the diagonal norm;
* face ``f`` of a triangle joins vertices (f, f+1); tet faces are
  (1,2,3),(1,4,2),(2,4,3),(1,3,4);
* ``uface[:,i] = sum_j interp[j,i] * u[:, perm[j,face]]``; the right element of
  an interface uses column ``nbrperm[i,orient]``;
* ``normal[:,face]`` is the reference outward normal scaled so that
  ``sum(wface)*|normal|`` is the face measure.

Indices stored here are 0-based (``index_base=0`` at the C ABI); all arrays are
Fortran-ordered so their memory layout equals the Julia arrays'.

p=1 Omega operators (3-/4-node) and the 6-node p=2 triangle operator are
unique; the 11-node p=2 tet and 12-node p=2 diagonal-E triangle operators are
built with a minimum-Frobenius-norm skew part, so their *values* are valid SBP
operators but are NOT claimed equal to SummationByParts.jl's ("parity
unpinned" for operator values).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from math import comb, factorial

import numpy as np

TRI_VTX = np.array([[-1.0, -1.0], [1.0, -1.0], [-1.0, 1.0]])
TET_VTX = np.array([[-1.0, -1.0, -1.0], [1.0, -1.0, -1.0],
                    [-1.0, 1.0, -1.0], [-1.0, -1.0, 1.0]])
# local vertex lists of each face (0-based)
TRI_FACE_VTX = np.array([[0, 1], [1, 2], [2, 0]])
TET_FACE_VTX = np.array([[0, 1, 2], [0, 3, 1], [1, 3, 2], [0, 2, 3]])


def _F(a):
    return np.asfortranarray(a)


# ----------------------------------------------------------------------------
# symmetric node orbits (barycentric) and cubature rules
# ----------------------------------------------------------------------------
def _orbit(bary):
    """All distinct permutations of a barycentric tuple, deterministic order."""
    return np.array(sorted(set(itertools.permutations(tuple(bary))), reverse=True))


def _ref_moment(exps):
    """Exact integral of prod x_d^e_d over the reference simplex."""
    dim = len(exps)
    tot = 0.0
    for sub in itertools.product(*[range(e + 1) for e in exps]):
        c = 1.0
        for e, a in zip(exps, sub):
            c *= comb(e, a) * 2.0 ** a * (-1.0) ** (e - a)
        num = 1.0
        for a in sub:
            num *= factorial(a)
        tot += c * num / factorial(sum(sub) + dim)
    return 2.0 ** dim * tot


def _exponents(dim, degree):
    return [e for e in itertools.product(range(degree + 1), repeat=dim)
            if sum(e) <= degree]


def vandermonde(x, degree):
    """Monomial basis at points x[dim, n] -> V[n, nb], and its derivatives."""
    dim, n = x.shape
    ex = _exponents(dim, degree)
    V = np.ones((n, len(ex)))
    dV = np.zeros((dim, n, len(ex)))
    for b, e in enumerate(ex):
        for d in range(dim):
            V[:, b] *= x[d] ** e[d]
        for d in range(dim):
            if e[d] == 0:
                continue
            t = e[d] * x[d] ** (e[d] - 1)
            for d2 in range(dim):
                if d2 != d:
                    t = t * x[d2] ** e[d2]
            dV[d, :, b] = t
    return V, dV


def cubature_error(x, w, degree):
    dim = x.shape[0]
    err = 0.0
    for e in _exponents(dim, degree):
        val = np.sum(w * np.prod([x[d] ** e[d] for d in range(dim)], axis=0))
        err = max(err, abs(val - _ref_moment(e)))
    return err


def _solve_weights(orbits_xy, degree):
    """Least-squares orbit weights making the rule exact to ``degree``."""
    dim = orbits_xy[0].shape[0]
    ex = _exponents(dim, degree)
    A = np.array([[np.sum(np.prod([X[d] ** e[d] for d in range(dim)], axis=0))
                   for X in orbits_xy] for e in ex])
    b = np.array([_ref_moment(e) for e in ex])
    w, *_ = np.linalg.lstsq(A, b, rcond=None)
    assert np.linalg.norm(A @ w - b) < 1e-13, "cubature weights do not exist"
    return w


def line_gauss_legendre(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return x, w


def line_lgl4():
    a = 1.0 / np.sqrt(5.0)
    return np.array([-1.0, -a, a, 1.0]), np.array([1.0, 5.0, 5.0, 1.0]) / 6.0


def tri_cubature(kind, degree):
    """(bary[n,3], w[n]) on the reference triangle (area 2)."""
    if kind == "omega" and degree == 1:
        b = _orbit((2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0))
        return b, np.full(3, 2.0 / 3.0)
    if kind == "omega" and degree == 2:
        # 6-point degree-4 interior rule (two S21 orbits); weights re-solved
        a1, a2 = 0.445948490915965, 0.091576213509771
        o1, o2 = _orbit((1 - 2 * a1, a1, a1)), _orbit((1 - 2 * a2, a2, a2))
        w = _solve_weights([(o1 @ TRI_VTX).T, (o2 @ TRI_VTX).T], 3)
        return np.vstack([o1, o2]), np.r_[np.full(3, w[0]), np.full(3, w[1])]
    if kind == "diage" and degree == 2:
        # vertices + LGL-4 edge nodes + one interior S21 orbit; degree 4.
        g = (1.0 - 1.0 / np.sqrt(5.0)) / 2.0
        ov = _orbit((1.0, 0.0, 0.0))
        oe = _orbit((1.0 - g, g, 0.0))
        a = 0.21285436  # Newton-polished below
        ex4 = [e for e in _exponents(2, 4)]

        def resid(p):
            os_ = _orbit((1 - 2 * p[3], p[3], p[3]))
            X = [(ov @ TRI_VTX).T, (oe @ TRI_VTX).T, (os_ @ TRI_VTX).T]
            return np.array([sum(p[k] * np.sum(X[k][0] ** e[0] * X[k][1] ** e[1])
                                 for k in range(3)) - _ref_moment(e) for e in ex4])

        p = np.array([0.02504451, 0.1072952, 0.42703176, a])
        for _ in range(20):
            r = resid(p)
            J = np.array([(resid(p + 1e-7 * np.eye(4)[k]) - r) / 1e-7
                          for k in range(4)]).T
            dp = np.linalg.lstsq(J, -r, rcond=None)[0]
            p = p + dp
            if np.linalg.norm(dp) < 1e-15:
                break
        os_ = _orbit((1 - 2 * p[3], p[3], p[3]))
        w = _solve_weights([(ov @ TRI_VTX).T, (oe @ TRI_VTX).T, (os_ @ TRI_VTX).T], 4)
        return (np.vstack([ov, oe, os_]),
                np.r_[np.full(3, w[0]), np.full(6, w[1]), np.full(3, w[2])])
    if kind == "omega_alt" and degree == 2:
        # 7 interior nodes (centroid + two S21 orbits), degree 3: a valid p=2 operator whose node count has no tuned
        # kernel instantiation (exercises the size-generic kernels; SummationByParts' SBP-Gamma p=2 also has 7 nodes)
        a1, a2 = 0.11, 0.46
        oc = _orbit((1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0))
        o1, o2 = _orbit((1 - 2 * a1, a1, a1)), _orbit((1 - 2 * a2, a2, a2))
        w = _solve_weights([(o @ TRI_VTX).T for o in (oc, o1, o2)], 3)
        assert w.min() > 0
        return np.vstack([oc, o1, o2]), np.r_[np.full(1, w[0]), np.full(3, w[1]), np.full(3, w[2])]
    raise ValueError(f"no triangle cubature for {kind} p={degree}")


def tet_cubature(kind, degree):
    """(bary[n,4], w[n]) on the reference tet (volume 4/3)."""
    if kind == "omega" and degree == 1:
        b_ = (5.0 - np.sqrt(5.0)) / 20.0
        b = _orbit((1 - 3 * b_, b_, b_, b_))
        return b, np.full(4, 1.0 / 3.0)
    if kind == "omega" and degree == 2:
        # centroid + S31 + S22, 11 interior nodes, positive weights, degree 3
        # (=2p-1, all SBP needs).  The degree-4 rule with this symmetry has a
        # negative centroid weight (Keast), which cannot serve as a norm.
        a1, a2 = 0.075, 0.415
        oc = _orbit((0.25, 0.25, 0.25, 0.25))
        o1 = _orbit((1 - 3 * a1, a1, a1, a1))
        o2 = _orbit((a2, a2, 0.5 - a2, 0.5 - a2))
        w = _solve_weights([(o @ TET_VTX).T for o in (oc, o1, o2)], 3)
        assert w.min() > 0
        return (np.vstack([oc, o1, o2]),
                np.r_[np.full(1, w[0]), np.full(4, w[1]), np.full(6, w[2])])
    if kind == "omega_alt" and degree == 2:
        # 14 interior nodes (two S31 orbits + S22), degree 3: no tuned instantiation for this size
        a1, a2, a3 = 0.09, 0.31, 0.42
        o1 = _orbit((1 - 3 * a1, a1, a1, a1))
        o2 = _orbit((1 - 3 * a2, a2, a2, a2))
        o3 = _orbit((a3, a3, 0.5 - a3, 0.5 - a3))
        w = _solve_weights([(o @ TET_VTX).T for o in (o1, o2, o3)], 3)
        assert w.min() > 0
        return np.vstack([o1, o2, o3]), np.r_[np.full(4, w[0]), np.full(4, w[1]), np.full(6, w[2])]
    if kind == "diage" and degree == 1:
        # SBPDiagonalE on the tet, p=1 (the reference's getTetSBPDiagE, solver/common.jl:306): the face nodes -- the 3-point
        # degree-2 rule of every face -- ARE volume nodes (12), plus the centroid; volume rule of degree 2p-1 = 1
        oc = _orbit((0.25, 0.25, 0.25, 0.25))
        of = _orbit((2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0, 0.0))
        w = _solve_weights([(o @ TET_VTX).T for o in (oc, of)], 1)
        assert w.min() > 0
        return np.vstack([oc, of]), np.r_[np.full(1, w[0]), np.full(12, w[1])]
    raise ValueError(f"no tet cubature for {kind} p={degree}")


# ----------------------------------------------------------------------------
# operator containers
# ----------------------------------------------------------------------------
@dataclass
class SBPFace:
    """Mirror of ``mesh.sbpface`` (SummationByParts TriFace/TetFace/SparseFace)."""
    numnodes: int            # nfn
    stencilsize: int         # ss (== sbp.numnodes for dense faces, 1 for sparse)
    interp: np.ndarray       # [ss, nfn]
    perm: np.ndarray         # [ss, numfaces] (sparse: [nfn, numfaces]) 0-based
    nbrperm: np.ndarray      # [nfn, norient] 0-based
    wface: np.ndarray        # [nfn]
    normal: np.ndarray       # [dim, numfaces]
    sparse: bool = False
    facenodes_bary: np.ndarray = field(default=None, repr=False)  # [nfn, dim]


@dataclass
class SBPOperator:
    """Mirror of ``sbp`` (AbstractOperator): numnodes, Q, w + node locations."""
    dim: int
    degree: int
    kind: str                # "omega" | "diage" | "gamma"
    numnodes: int
    Q: np.ndarray            # [nn, nn, dim]
    w: np.ndarray            # [nn]
    bary: np.ndarray         # [nn, dim+1] barycentric node coordinates
    xref: np.ndarray         # [dim, nn] reference coordinates
    face: SBPFace = None
    E: np.ndarray = None     # [nn, nn, dim] boundary operators

    @property
    def numfaces(self):
        return self.dim + 1


def _find_perm(xa, xb, tol=1e-12):
    """perm with xb[:, perm[j]] == xa[:, j]."""
    perm = np.empty(xa.shape[1], dtype=np.int64)
    for j in range(xa.shape[1]):
        d = np.linalg.norm(xb - xa[:, j:j + 1], axis=0)
        k = int(np.argmin(d))
        assert d[k] < tol, "node set is not symmetric"
        perm[j] = k
    return perm


def _min_norm_skew(nn, V, rhs):
    """Skew-symmetric S of minimum Frobenius norm with S @ V == rhs."""
    pairs = [(i, j) for i in range(nn) for j in range(i + 1, nn)]
    nb = V.shape[1]
    A = np.zeros((nn * nb, len(pairs)))
    for c, (i, j) in enumerate(pairs):
        # S[i,j] = s, S[j,i] = -s
        A[i * nb:(i + 1) * nb, c] += V[j]
        A[j * nb:(j + 1) * nb, c] -= V[i]
    s, *_ = np.linalg.lstsq(A, rhs.reshape(-1), rcond=None)
    assert np.linalg.norm(A @ s - rhs.reshape(-1)) < 1e-11, \
        "SBP accuracy conditions are not compatible with the cubature"
    S = np.zeros((nn, nn))
    for c, (i, j) in enumerate(pairs):
        S[i, j] = s[c]
        S[j, i] = -s[c]
    return S


def build_operator(dim: int, degree: int, kind: str = "omega") -> SBPOperator:
    """Construct a degree-``degree`` SBP operator and its face operator.

    kind: "omega" (interior nodes, dense face interpolation; the reference's
    default ``operator_type=SBPOmega``), "diage" (SBPDiagonalE, face nodes
    coincide with volume nodes, sparse face), "gamma" (p=1 vertex nodes with a dense
    Gauss face operator; only used to check the golden volume blocks of
    test_lowlevel.jl:766-807, which were produced with SBPGamma).
    """
    vtx = TRI_VTX if dim == 2 else TET_VTX
    fvtx = TRI_FACE_VTX if dim == 2 else TET_FACE_VTX
    nfaces = dim + 1
    if kind == "gamma":
        assert degree == 1
        bary = np.eye(dim + 1)
        vol = 2.0 if dim == 2 else 4.0 / 3.0
        w = np.full(dim + 1, vol / (dim + 1))
    elif dim == 2:
        bary, w = tri_cubature(kind, degree)
    else:
        bary, w = tet_cubature(kind, degree)
    nn = bary.shape[0]
    x = (bary @ vtx).T.copy()                   # [dim, nn]
    V, dV = vandermonde(x, degree)

    # ---- face cubature in barycentric coordinates of the face -------------
    if dim == 2:
        if kind == "diage":
            t, wf = line_lgl4()
        else:
            t, wf = line_gauss_legendre(degree + 1)
        fb = np.stack([(1 - t) / 2, (1 + t) / 2], axis=1)       # [nfn, 2]
        ref_n = np.array([[0.0, -1.0], [1.0, 1.0], [-1.0, 0.0]]).T
    else:
        fb, wf = tri_cubature("omega", degree)
        wf = wf * 1.0
        ref_n = np.array([[0.0, 0.0, -1.0], [0.0, -1.0, 0.0],
                          [1.0, 1.0, 1.0], [-1.0, 0.0, 0.0]]).T
    nfn = fb.shape[0]

    # face node coordinates on each face, and the per-face interpolation
    xf = [(fb @ vtx[fvtx[f]]).T for f in range(nfaces)]         # [dim, nfn]
    Vinv = np.linalg.pinv(V)
    R = [vandermonde(xf[f], degree)[0] @ Vinv for f in range(nfaces)]  # [nfn, nn]

    sparse = kind == "diage"
    if sparse:
        perm = np.zeros((nfn, nfaces), dtype=np.int64)
        for f in range(nfaces):
            perm[:, f] = _find_perm(xf[f], x)
            Rf = np.zeros((nfn, nn))
            Rf[np.arange(nfn), perm[:, f]] = 1.0
            R[f] = Rf
        interp = np.ones((1, nfn))
        ss = 1
    else:
        # one interp matrix shared by all faces through the symmetry map that
        # carries face 0 (and its opposite vertex) onto face f
        perm = np.zeros((nn, nfaces), dtype=np.int64)
        for f in range(nfaces):
            opp0 = [v for v in range(dim + 1) if v not in fvtx[0]][0]
            oppf = [v for v in range(dim + 1) if v not in fvtx[f]][0]
            src = list(fvtx[0]) + [opp0]
            dst = list(fvtx[f]) + [oppf]
            vmap = np.zeros(dim + 1, dtype=int)
            vmap[src] = dst
            bmapped = np.zeros_like(bary)
            bmapped[:, vmap] = bary            # barycentric weight of vertex v moves to vmap[v]
            perm[:, f] = _find_perm((bmapped @ vtx).T, x)
            assert np.allclose(R[f][:, perm[:, f]], R[0], atol=1e-12)
        interp = R[0].T.copy()                 # [nn, nfn]
        ss = nn

    # nbrperm: face node i of elementL coincides with node nbrperm[i,o] of elementR
    if dim == 2:
        nbr = np.zeros((nfn, 1), dtype=np.int64)
        nbr[:, 0] = _find_perm(fb[:, ::-1].T, fb.T)
    else:
        nbr = np.zeros((nfn, 3), dtype=np.int64)
        # elementR lists the shared vertices as (l1,l3,l2), (l3,l2,l1), (l2,l1,l3)
        for o, sig in enumerate([(0, 2, 1), (2, 1, 0), (1, 0, 2)]):
            nbr[:, o] = _find_perm(fb[:, list(sig)].T, fb.T)

    # ---- boundary operators and Q ------------------------------------------
    E = np.zeros((nn, nn, dim))
    for f in range(nfaces):
        for d in range(dim):
            E[:, :, d] += R[f].T @ np.diag(wf * ref_n[d, f]) @ R[f]
    H = np.diag(w)
    Q = np.zeros((nn, nn, dim))
    for d in range(dim):
        rhs = H @ dV[d] - 0.5 * E[:, :, d] @ V
        S = _min_norm_skew(nn, V, rhs)
        Q[:, :, d] = S + 0.5 * E[:, :, d]

    face = SBPFace(numnodes=nfn, stencilsize=ss, interp=_F(interp), perm=_F(perm),
                   nbrperm=_F(nbr), wface=wf.copy(), normal=_F(ref_n),
                   sparse=sparse, facenodes_bary=fb)
    op = SBPOperator(dim=dim, degree=degree, kind=kind, numnodes=nn, Q=_F(Q),
                     w=w.copy(), bary=bary, xref=_F(x), face=face, E=_F(E))
    check_operator(op)
    return op


def face_matrices(op: SBPOperator):
    """Dense per-face interpolation matrices R_f [nfn, nn] (perm folded in)."""
    fo = op.face
    nn, nfn = op.numnodes, fo.numnodes
    out = []
    for f in range(op.numfaces):
        R = np.zeros((nfn, nn))
        if fo.sparse:
            R[np.arange(nfn), fo.perm[:, f]] = 1.0
        else:
            for j in range(fo.stencilsize):
                R[:, fo.perm[j, f]] += fo.interp[j, :]
        out.append(R)
    return out


def check_operator(op: SBPOperator, tol=1e-11):
    """SBP identities every operator (own or supplied) must satisfy
    (mirrors test/euler/test_curvilinear.jl:22-81)."""
    nn, dim, p = op.numnodes, op.dim, op.degree
    V, dV = vandermonde(np.asarray(op.xref), p)
    one = np.ones(nn)
    vol = 2.0 if dim == 2 else 4.0 / 3.0
    assert abs(op.w.sum() - vol) < tol, "sum(w) != |reference element|"
    assert op.w.min() > 0, "norm must be positive"
    R = face_matrices(op)
    for d in range(dim):
        Qd = op.Q[:, :, d]
        assert np.abs(Qd @ one).max() < tol, "Q*1 != 0"
        D = Qd / op.w[:, None]
        assert np.abs(D @ V - dV[d]).max() < 10 * tol, "D not exact on P_p"
        E = sum(R[f].T @ np.diag(op.face.wface * op.face.normal[d, f]) @ R[f]
                for f in range(op.numfaces))
        assert np.abs(Qd + Qd.T - E).max() < tol, "Q + Q^T != E"
        assert abs(one @ E @ one) < tol, "1^T E 1 != 0"
    for o in range(op.face.nbrperm.shape[1]):
        pr = op.face.nbrperm[:, o]
        assert np.array_equal(pr[pr], np.arange(op.face.numnodes)), \
            "nbrperm is not an involution"
    return True
