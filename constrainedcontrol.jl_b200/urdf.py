"""Minimal URDF ingestion: `Mechanism(path; floating=false, g)` for serial/branching trees of revolute, prismatic and
fixed joints (reference examples/lqr_sawyer.jl:8-9 loads examples/examples_files/sawyer_arm.urdf this way).

Maximal-coordinate convention: one body per link with its frame at the link's centre of mass, oriented like the
link's <inertial> frame; with floating=False the root link is merged into the origin.  Bodies and eqcs are created in
URDF *document order* (upstream keeps them in a Dict, whose order is hash dependent -- SURVEY.md 7 "URDF body ordering";
this order is the restatement-defined, deterministic one).
"""
from __future__ import annotations

import json
import xml.etree.ElementTree as ET

import numpy as np

from .mechanism import (Body, EqualityConstraint, Fixed, Mechanism, Origin, Prismatic, Revolute, propagate_positions, qconj, qmul, qrot,
                        quat_from_rpy)


def _floats(s, n, default=0.0):
    if s is None:
        return np.full(n, default)
    return np.array([float(v) for v in s.split()], dtype=np.float64)


def parse_urdf(path: str) -> dict:
    """-> plain dict {links: [{name, mass, com, rpy, inertia(3x3)}], joints: [{name, type, parent, child, xyz, rpy, axis}]}"""
    root = ET.parse(path).getroot()
    links, joints = [], []
    for l in root.findall("link"):
        ine = l.find("inertial")
        if ine is None:
            links.append(dict(name=l.get("name"), mass=0.0, com=[0, 0, 0], rpy=[0, 0, 0], inertia=np.zeros((3, 3)).tolist()))
            continue
        o = ine.find("origin")
        I = ine.find("inertia")
        J = np.array([[float(I.get("ixx")), float(I.get("ixy")), float(I.get("ixz"))],
                      [float(I.get("ixy")), float(I.get("iyy")), float(I.get("iyz"))],
                      [float(I.get("ixz")), float(I.get("iyz")), float(I.get("izz"))]])
        links.append(dict(name=l.get("name"), mass=float(ine.find("mass").get("value")),
                          com=_floats(o.get("xyz") if o is not None else None, 3).tolist(),
                          rpy=_floats(o.get("rpy") if o is not None else None, 3).tolist(), inertia=J.tolist()))
    for j in root.findall("joint"):
        o = j.find("origin")
        ax = j.find("axis")
        joints.append(dict(name=j.get("name"), type=j.get("type"), parent=j.find("parent").get("link"), child=j.find("child").get("link"),
                           xyz=_floats(o.get("xyz") if o is not None else None, 3).tolist(),
                           rpy=_floats(o.get("rpy") if o is not None else None, 3).tolist(),
                           axis=_floats(ax.get("xyz") if ax is not None else "1 0 0", 3).tolist()))
    return dict(links=links, joints=joints)


def drop_massless_frames(desc: dict) -> dict:
    """Links without an <inertial> element (camera / wrist / mounting FRAMES such as right_hand_camera, right_wrist, head_camera,
    controller_box, pedestal_feet of examples/examples_files/sawyer.urdf) that hang on a fixed joint and carry no further links are not
    bodies: with m = 0 and J = 0 their 6x6 dynamics block is identically zero and the maximal-coordinate KKT system is singular (in the
    reference as much as here).  The same holds for the file's placeholder links (right_hand, right_l1_2, right_l2_2, right_l4_2: m = 1e-8 kg
    and an inertia "tensor" with all nine entries 1e-8 -- rank 1, not a physical inertia: the rotational pivot block of the tree-ordered
    LDU is singular).  Both kinds are removed together with their fixed joint (their 1e-8 kg are below the rounding of the 1-5 kg links they
    are welded to); such a link on a movable joint, or with links depending on it, is an error."""
    links = list(desc["links"]); joints = list(desc["joints"])
    changed = True
    while changed:
        changed = False
        parents = {j["parent"] for j in joints}
        for l in links:
            J = np.array(l["inertia"], dtype=float)
            ev = np.linalg.eigvalsh(0.5 * (J + J.T))
            massless = (l["mass"] == 0.0 and not np.any(J)) or (np.any(J) and ev.min() <= 1e-9 * ev.max())
            jin = [j for j in joints if j["child"] == l["name"]]
            if massless and jin and l["name"] not in parents:
                if any(j["type"] != "fixed" for j in jin):
                    raise ValueError(f"URDF link {l['name']} has no inertia but hangs on a movable joint")
                links = [x for x in links if x is not l]
                joints = [j for j in joints if j["child"] != l["name"]]
                changed = True
                break
    return dict(links=links, joints=joints)


def mechanism_from_description(desc: dict, floating: bool = False, g: float = -9.81, dt: float = 0.01) -> Mechanism:
    if floating:
        raise NotImplementedError("floating bases are outside the hot-path scope")
    desc = drop_massless_frames(desc)
    links = {l["name"]: l for l in desc["links"]}
    children = {j["child"] for j in desc["joints"]}
    roots = [l["name"] for l in desc["links"] if l["name"] not in children]
    assert len(roots) == 1, "URDF must have exactly one root link"
    root = roots[0]
    origin = Origin()
    bodies, body_of = [], {}
    for l in desc["links"]:
        if l["name"] == root:
            continue
        b = Body(l["mass"], np.array(l["inertia"]), name=l["name"])
        bodies.append(b)
        body_of[l["name"]] = b
    eqcs = []
    for j in desc["joints"]:
        lp, lc = links[j["parent"]], links[j["child"]]
        child = body_of[j["child"]]
        q_j = quat_from_rpy(*j["rpy"])
        q_ic = quat_from_rpy(*lc["rpy"])
        if j["parent"] == root:
            parent, q_ip, com_p = origin, np.array([1.0, 0, 0, 0]), np.zeros(3)
        else:
            parent, q_ip, com_p = body_of[j["parent"]], quat_from_rpy(*lp["rpy"]), np.array(lp["com"])
        p1 = qrot(qconj(q_ip), np.array(j["xyz"]) - com_p)          # joint vertex in the parent body frame
        p2 = qrot(qconj(q_ic), -np.array(lc["com"]))                # ... and in the child body frame
        q_pj = qmul(qconj(q_ip), q_j)                               # parent body frame -> joint frame
        qoff = qmul(q_pj, q_ic)
        axis = qrot(q_pj, np.array(j["axis"]))                      # URDF axis is given in the joint (= child link) frame
        if j["type"] in ("revolute", "continuous"):
            jt = Revolute(parent, child, axis, p1=p1, p2=p2, qoffset=qoff)
        elif j["type"] == "prismatic":
            jt = Prismatic(parent, child, axis, p1=p1, p2=p2, qoffset=qoff)
        elif j["type"] == "fixed":
            jt = Fixed(parent, child, p1=p1, p2=p2, qoffset=qoff)
        else:
            raise NotImplementedError(f"URDF joint type {j['type']}")
        eqcs.append(EqualityConstraint(jt, name=j["name"]))
    mech = Mechanism(origin, bodies, eqcs, g=g, dt=dt)
    zero_configuration(mech)
    return mech


def zero_configuration(mech: Mechanism, coords=None):
    """place every body by forward kinematics from the joint coordinates (all zero by default)"""
    coords = coords or {}
    pending = list(range(mech.ne))
    placed = set()
    guard = 0
    while pending and guard < 100 * (mech.ne + 1):
        guard += 1
        i = pending.pop(0)
        e = mech.eqconstraints[i]
        j = e.constraints[0]
        if not isinstance(j.parent, Origin) and id(j.parent) not in placed:
            pending.append(i)
            continue
        if j.nT == 3 and j.nR == 3:
            qb = qmul(j.parent.state.qc, j.qoffset)
            j.child.state.qc = qb
            j.child.state.xc = j.parent.state.xc + qrot(j.parent.state.qc, j.p1) - qrot(qb, j.p2)
        else:
            from .mechanism import setJointPosition
            setJointPosition(mech, e, [coords.get(i, 0.0)])
        placed.add(id(j.child))
    assert not pending, "URDF joints do not form a tree rooted at the base link"


def mechanism_from_urdf(path: str, floating: bool = False, g: float = -9.81, dt: float = 0.01) -> Mechanism:
    """Mechanism(path, floating=false, g=...)"""
    return mechanism_from_description(parse_urdf(path), floating=floating, g=g, dt=dt)


def mechanism_from_json(path: str, **kw) -> Mechanism:
    return mechanism_from_description(json.load(open(path)), **kw)
