"""Host-side initial conditions (the reference evaluates ICs on the host too:
``src/solver/euler/ic.jl`` loops ``calc<Name>(params, coords[:,j,i], q[:,j,i])``
over the mesh).  Vectorised numpy restatements of the three exact solutions the
scoped configurations use (``common_funcs.jl:25-78, 204-283, 312-351, 841-857,
899-936``); the CPU oracle holds the scalar versions the tests compare with.
"""
from __future__ import annotations

import numpy as np


def ICIsentropicVortex(coords, params):
    g, R = params.gamma, params.R
    cv = R / (g - 1)
    dim = coords.shape[0]
    x, y = coords[0], coords[1]
    theta3 = None
    if dim == 3:
        z = coords[2]
        phi_z = np.pi / 4
        theta1 = np.arctan2(z, x)
        phi2 = 0.5 * np.pi - theta1
        r_xz = np.sqrt(x * x + z * z)
        x = r_xz * np.sin(phi_z + phi2)
        theta3 = theta1 + phi_z + phi2 - 0.5 * np.pi
        theta = np.arctan2(x, y)
    else:
        theta = np.arctan2(y, x)
    r_in, rho_in, M_in, p_in = 1.0, 2.0, 0.95, 1 / g
    r = np.sqrt(x * x + y * y)
    tmp1 = ((g - 1) / 2) * M_in * M_in
    rho_r = rho_in * (1 + tmp1 * (1 - (r_in * r_in) / (r * r))) ** (1 / (g - 1))
    p_r = p_in * (rho_r / rho_in) ** g
    a_r = np.sqrt(g * p_r / rho_r)
    M_r = np.sqrt((2 / (g - 1)) * ((rho_in / rho_r) ** (g - 1)) * (1 + tmp1) - 2 / (g - 1))
    U_r = M_r * a_r
    e_r = cv * p_r / (rho_r * R)
    E_r = rho_r * e_r + 0.5 * rho_r * U_r * U_r
    q = np.empty((dim + 2,) + coords.shape[1:], order="F")
    q[0] = rho_r
    if dim == 2:
        q[1] = rho_r * (U_r * np.sin(theta))
        q[2] = rho_r * (-U_r * np.cos(theta))
    else:
        v_r = U_r * np.sin(theta)
        u_r = -U_r * np.cos(theta)
        q[1] = rho_r * (u_r * np.cos(theta3))
        q[2] = rho_r * v_r
        q[3] = rho_r * (u_r * np.sin(theta3))
    q[dim + 1] = E_r
    return q


def ICExp(coords, params):
    g1 = params.gamma - 1
    dim = coords.shape[0]
    q = np.empty((dim + 2,) + coords.shape[1:], order="F")
    if dim == 2:
        xy = coords[0] * coords[1]
        af, b = 1.0 / 5, 0.01
        q[0] = np.exp(af * xy + b)
        q[1] = np.exp(af * 2 * xy + b)
        q[2] = np.exp(af * 3 * xy + b)
        q[3] = (1 / g1 + 0.5) * np.exp(af * 5 * xy + b) + 0.5 * np.exp(af * 3 * xy + b)
    else:
        a, b = 1.0 / 500, 0.01
        c1, c2, c3, c4, c5 = 1, 2, 3, 4, 20
        d1, d2, d3, d4, d5 = 1, 0.05, 0.15, 0.25, 1
        xyz = coords[0] * coords[1] * coords[2]
        t2 = np.exp(b)
        t3 = a * c1 * xyz
        q[0] = d1 * t2 * np.exp(t3)
        q[1] = d2 * t2 * np.exp(a * c2 * xyz)
        q[2] = d3 * t2 * np.exp(a * c3 * xyz)
        q[3] = d4 * t2 * np.exp(a * c4 * xyz)
        q[4] = (t2 * np.exp(-t3) * (d2 * d2 * np.exp(a * c2 * xyz * 2.0) + d3 * d3 * np.exp(a * c3 * xyz * 2.0)
                                    + d4 * d4 * np.exp(a * c4 * xyz * 2.0)) * 0.5) / d1 \
            + (d5 * t2 * np.exp(a * c5 * xyz)) / g1
    return q


def ICFreeStream(coords, params):
    dim = coords.shape[0]
    q = np.empty((dim + 2,) + coords.shape[1:], order="F")
    q[0] = params.rho_free
    q[1] = params.rho_free * params.Ma * np.cos(params.aoa)
    if dim == 2:
        q[2] = params.rho_free * params.Ma * np.sin(params.aoa)
    else:
        q[2] = 0.0
        q[3] = -params.rho_free * params.Ma * np.sin(params.aoa)
    q[dim + 1] = params.E_free
    return q


ICDict = {"ICIsentropicVortex": ICIsentropicVortex, "ICExp": ICExp, "ICFreeStream": ICFreeStream}
