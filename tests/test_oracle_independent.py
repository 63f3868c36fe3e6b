"""An INDEPENDENT numpy transcription of the published step equations (papers at reference README.md:14,16),
written without looking at oracle/*.hpp, evaluated at the oracle's solution of one integrator step.

It shares no code with the oracle: constraint functions are typed in from A.3, their Jacobians are taken by central finite differences in
the tangent coordinates (dx, phi) with q <- q (x) [1; phi], and the multipliers are ELIMINATED (the dynamics residual without constraint
forces must lie in the row space of G), so the check is independent of the oracle's choice of constraint-row basis and lambda scaling."""
import numpy as np
import pytest


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1:] @ b[1:], *(a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:]))])


def qconj(a):
    return np.array([a[0], -a[1], -a[2], -a[3]])


def rot(q, p):
    return qmul(qmul(q, np.array([0.0, *p])), qconj(q))[1:]


def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def basis_perp(ax):
    ax = ax / np.linalg.norm(ax)
    h = np.eye(3)[np.argmin(np.abs(ax))]
    v1 = np.cross(ax, h); v1 /= np.linalg.norm(v1)
    return np.stack([v1, np.cross(ax, v1)])


def joint_g(j, xa, qa, xb, qb):
    """A.3: translational part R_a^T (x_b + R_b p2 - x_a) - p1, rotational part vec(q_a^-1 q_b q_off^-1); nT/nR of 3, 2 (perp. to axis) rows"""
    gT = rot(qconj(qa), xb + rot(qb, j.p2) - xa) - j.p1
    gR = qmul(qmul(qconj(qa), qb), qconj(j.qoffset))[1:]
    V = basis_perp(j.axis)
    out = []
    for n, g in ((j.nT, gT), (j.nR, gR)):
        if n == 3: out.append(g)
        elif n == 2: out.append(V @ g)
        elif n == 1: out.append(np.array([j.axis @ g / np.linalg.norm(j.axis)]))
    return np.concatenate(out)


def all_g(mech, X, Q):
    out = []
    for e in mech.eqconstraints:
        for j in e.constraints:
            a = None if j.parent is mech.origin else mech.bodies.index(j.parent)
            b = mech.bodies.index(j.child)
            xa, qa = (np.zeros(3), np.array([1.0, 0, 0, 0])) if a is None else (X[a], Q[a])
            out.append(joint_g(j, xa, qa, X[b], Q[b]))
    return np.concatenate(out)


def tangent_jacobian(mech, X, Q, h=1e-6):
    """G = d g / d (dx_b, phi_b) for every body, central differences; q <- q (x) [1; phi] (A.2)"""
    nb = mech.nb
    g0 = all_g(mech, X, Q)
    G = np.zeros((g0.size, 6 * nb))
    for b in range(nb):
        for k in range(6):
            cols = []
            for sgn in (+1, -1):
                Xp = [x.copy() for x in X]; Qp = [q.copy() for q in Q]
                if k < 3:
                    Xp[b][k] += sgn * h
                else:
                    phi = np.zeros(3); phi[k - 3] = sgn * h
                    Qp[b] = qmul(Q[b], np.array([1.0, *phi]))
                cols.append(all_g(mech, Xp, Qp))
            G[:, 6 * b + k] = (cols[0] - cols[1]) / (2 * h)
    return G


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "acrobot", "triple_cartpole"])
def test_oracle_step_satisfies_the_published_equations(cc, orc, name):
    mech, cfg = getattr(cc.examples, name)()
    rng = np.random.default_rng(5)
    # a generic state on the constraint manifold with consistent velocities
    th = rng.uniform(-0.7, 0.7, mech.ne); rate = rng.uniform(-1.0, 1.0, mech.ne)
    cc.mechanism.propagate_minimal(mech, dict(enumerate(th)), dict(enumerate(rate)))
    st0 = mech.state_vector()[None].copy()
    om = orc.OracleMech(mech.desc())
    dt, grav = mech.dt, mech.g
    # two steps: the first brings the knot onto the manifold exactly as the integrator defines it; check the SECOND step
    r = om.simulate(st0, 2, record=True)
    s1, s2 = r["storage"][1, 0], r["states"][0]          # continuous states at the start of step 2 and after it
    nb = mech.nb
    c4 = 4.0 / dt ** 2
    X2, Q2, X3, Q3, dfree = [], [], [], [], []
    for b, body in enumerate(mech.bodies):
        xc, qc, vc, wc = s1[b, 0:3], s1[b, 3:7], s1[b, 7:10], s1[b, 10:13]
        v2, w2 = s2[b, 7:10], s2[b, 10:13]               # solution of step 2
        x2 = xc + vc * dt                                # knot (discretizestate!)
        q2 = qmul(qc, dt / 2 * np.array([np.sqrt(c4 - wc @ wc), *wc]))
        x3 = x2 + v2 * dt
        q3 = qmul(q2, dt / 2 * np.array([np.sqrt(c4 - w2 @ w2), *w2]))
        assert np.allclose(s2[b, 0:3], x2, atol=1e-13) and np.allclose(s2[b, 3:7], q2, atol=1e-13)   # updatestate!: (xc,qc) <- (x2,q2)
        J, m = body.J, body.m
        dT = m * ((v2 - vc) / dt + np.array([0, 0, -grav]))
        dR = (np.sqrt(c4 - w2 @ w2) * np.eye(3) + skew(w2)) @ J @ w2 - (np.sqrt(c4 - wc @ wc) * np.eye(3) - skew(wc)) @ J @ wc
        X2.append(x2); Q2.append(q2); X3.append(x3); Q3.append(q3); dfree.append(np.concatenate([dT, dR]))
    dfree = np.concatenate(dfree)
    # (1) position-level constraints hold at the next knot
    assert np.abs(all_g(mech, X3, Q3)).max() < 1e-10
    # (2) the unconstrained dynamics residual is a combination of the constraint Jacobian rows at the CURRENT knot: d_free = G^T lambda
    G = tangent_jacobian(mech, X2, Q2)
    lam, res, rank, _ = np.linalg.lstsq(G.T, dfree, rcond=None)
    assert rank == G.shape[0]
    assert np.linalg.norm(G.T @ lam - dfree) < 1e-6 * max(1.0, np.linalg.norm(dfree))
    # ... and a wrong time level (Jacobian at the NEXT knot) or a dropped gravity term does NOT satisfy it: the check has teeth
    G3 = tangent_jacobian(mech, X3, Q3)
    lam3 = np.linalg.lstsq(G3.T, dfree, rcond=None)[0]
    assert np.linalg.norm(G3.T @ lam3 - dfree) > 1e-4 * np.linalg.norm(dfree)
