"""Host-side logic that needs no device: config dimensions, Storage views, weight assembly, batch sharding."""
import math
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_config_dimensions_match_survey_appendix_b(cc):
    # config -> (Nb, Ne, m_lambda, n = 12 Nb, mu)
    table = {
        "pendulum": (1, 1, 5, 12, 1), "cartpole": (2, 2, 10, 24, 1), "triple_cartpole": (4, 4, 20, 48, 1), "deltabot": (5, 5, 33, 60, 2),
        "doublependulum": (2, 2, 10, 24, 2), "acrobot": (2, 2, 10, 24, 1), "prismatic": (1, 1, 5, 12, 1),
    }
    for name, (nb, ne, ml, n, mu) in table.items():
        mech, cfg = getattr(cc.examples, name)()
        assert (mech.nb, mech.ne, mech.ml, 12 * mech.nb, len(cfg["eqcids"])) == (nb, ne, ml, n, mu), name
        assert np.abs(mech.constraint_values()).max() < 1e-12, name           # every example starts on its constraint manifold
    for n_links, dims in ((8, (9, 9, 45, 108, 1)), (32, (33, 33, 165, 396, 1))):
        mech, cfg = cc.examples.cartpole_n_pendulum(n_links)
        assert (mech.nb, mech.ne, mech.ml, 12 * mech.nb, len(cfg["eqcids"])) == dims
    mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
    assert (mech.nb, mech.ne, mech.ml, 12 * mech.nb, len(cfg["eqcids"])) == (7, 7, 35, 84, 7)
    # deltabot: one eqc with three joints (Revolute + Revolute + FixedOrientation = 13 rows), examples/lqr_deltabot.jl:25
    mech, _ = cc.examples.deltabot()
    assert sorted(len(e) for e in mech.eqconstraints) == [5, 5, 5, 5, 13]


def test_ids_eqcs_first_and_lookup(cc):
    mech, cfg = cc.examples.cartpole()
    assert cc.getid(mech.eqconstraints) == [1, 2] and cc.getid(mech.bodies) == [3, 4]      # eqcs 1..Ne, bodies Ne+1..Ne+Nb (SURVEY A.1 decision)
    assert mech.eqc_index(cfg["eqcids"][0]) == 0 and mech.body_index(4) == 1
    assert cc.geteqconstraint(mech, "joint2").id == 2
    assert mech.is_1dof(0) and mech.is_1dof(1)
    d, _ = cc.examples.deltabot()
    assert [d.is_1dof(i) for i in range(5)] == [True, True, False, True, True]


def test_storage_views_are_body_major(cc):
    st = cc.Storage(4, 2, B=3)
    st.data[...] = np.arange(st.data.size).reshape(st.data.shape)
    assert len(st) == 4 and len(st.x) == 2
    assert np.array_equal(st.x[1][2], st.data[2, 0, 1, 0:3]) and np.array_equal(st.q[0][3], st.data[3, 0, 0, 3:7])
    assert np.array_equal(st.v[1][0], st.data[0, 0, 1, 7:10]) and np.array_equal(st.w[0][1], st.data[1, 0, 0, 10:13])


def test_blockdiag_is_cat_dims_1_2(cc):
    Q = cc.blockdiag([np.eye(2) * 2, np.ones((1, 1)) * 5])
    assert Q.shape == (3, 3) and np.array_equal(Q, np.array([[2, 0, 0], [0, 2, 0], [0, 0, 5.0]]))


@pytest.mark.parametrize("total,world", [(4096, 8), (1024, 8), (10, 3), (7, 8), (1, 1)])
def test_shard_ranges_partition_the_batch(cc, total, world):
    r = [cc.dist.shard_range(total, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == total
    assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
    sizes = [hi - lo for lo, hi in r]
    assert max(sizes) - min(sizes) <= 1


def test_pid_constructor_scalar_form_and_asserts(cc):
    # pid.jl:16-23: PID(mechanism, eqcid::Integer, goal; P, I, D) wraps the scalars
    mech, cfg = cc.examples.pendulum()
    pid = cc.PID(mech, cfg["eqcids"][0], math.pi / 2, P=10.0, I=1.0, D=2.0)
    assert pid.goals.shape == (1, 1) and pid.P.tolist() == [10.0]
    d, _ = cc.examples.deltabot()
    with pytest.raises(AssertionError, match="Only 1 DOF"):
        cc.PID(d, d.eqconstraints[2].id, 0.0)


def test_are_util_batches_and_broadcasts_without_touching_the_device(cc):
    """util._batched (the host half of dare/care/lqr/dlqr): shared matrices broadcast over the batch, single problems stay single"""
    import numpy as np
    A = np.arange(8.0).reshape(2, 2, 2); B = np.ones((2, 1)); Q = np.eye(2); R = 2.0
    Ab, Bb, Qb, Rb, S, n, m, single = cc.util._batched(A, B, Q, np.asarray(R))
    assert (S, n, m, single) == (2, 2, 1, False)
    assert Ab.shape == (2, 2, 2) and Bb.shape == (2, 2, 1) and Qb.shape == (2, 2, 2) and Rb.shape == (2, 1, 1)
    assert Bb.flags.c_contiguous and np.array_equal(Bb[1], B) and Rb[1, 0, 0] == 2.0
    _, _, _, _, S, n, m, single = cc.util._batched(A[0], B[:, 0], Q, [[2.0]])
    assert (S, n, m, single) == (1, 2, 1, True)
    import pytest
    with pytest.raises(AssertionError):
        cc.util._batched(A, np.ones((3, 2, 1)), Q, [[1.0]])      # batch sizes differ


def test_deltabot_initial_condition_grid_matches_the_example(cc):
    """examples/lqr_deltabot.jl:57-145: the 101 x 101 grid has 1973 reachable platform positions; every initial
    condition built from it closes both kinematic loops"""
    import numpy as np
    A, al, ar = cc.examples.deltabot_initial_conditions()
    assert A.shape == (1973, 2) and (A[:, 1] >= 0).all()
    mech, _ = cc.examples.deltabot()
    nominal = mech.state_vector().copy()
    st = cc.examples.deltabot_states(mech, 2048)
    assert np.array_equal(mech.state_vector(), nominal)            # the helper leaves the mechanism where it found it
    worst = 0.0
    for s in st[::23]:
        mech.set_state_vector(s)
        worst = max(worst, float(np.abs(mech.constraint_values()).max()))
    mech.set_state_vector(nominal)
    assert worst < 1e-12
    assert np.unique(st.reshape(2048, -1), axis=0).shape[0] == 2048


def test_full_sawyer_urdf_ingestion_rules():
    """host logic of SURVEY 8(f) n3 (no GPU): the full sawyer.urdf description (tests/golden/sawyer_full.json, numbers parsed from the reference's
    file) loses its massless frames and its rank-1-inertia placeholder links, keeps document order, and rejects descriptions it cannot represent"""
    import json, os
    import numpy as np
    import ccb200 as cc
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sawyer_full.json")
    desc = json.load(open(gold))
    assert len(desc["links"]) == 24 and len(desc["joints"]) == 23
    kept = cc.urdf.drop_massless_frames(desc)
    dropped = {l["name"] for l in desc["links"]} - {l["name"] for l in kept["links"]}
    assert dropped == {"controller_box", "pedestal_feet", "right_hand_camera", "right_wrist", "head_camera", "right_hand", "right_l1_2", "right_l2_2", "right_l4_2"}
    assert [l["name"] for l in kept["links"]] == [l["name"] for l in desc["links"] if l["name"] not in dropped]        # document order kept
    assert all(j["type"] == "fixed" for j in desc["joints"] if j["child"] in dropped)
    mech = cc.urdf.mechanism_from_description(desc, floating=False, g=-9.81)
    assert (mech.nb, mech.ne, mech.ml) == (14, 14, 76)
    assert [e.name for e in mech.eqconstraints][:4] == ["pedestal_fixed", "right_arm_mount", "torso_t0", "right_j0"]
    assert abs(np.abs(mech.constraint_values()).max()) < 1e-12                 # the zero configuration satisfies every joint
    # a massless link on a movable joint cannot be represented in maximal coordinates: refused, not silently dropped
    bad = json.loads(json.dumps(desc))
    for j in bad["joints"]:
        if j["child"] == "right_wrist":
            j["type"] = "revolute"
    try:
        cc.urdf.drop_massless_frames(bad)
        assert False, "expected ValueError"
    except ValueError as e:
        assert "right_wrist" in str(e)
    # the descriptions used elsewhere are untouched by the rule
    for name, dims in (("sawyer_simple.json", (10, 10, 52)), ("sawyer_arm.json", (7, 7, 35))):
        m = cc.urdf.mechanism_from_json(os.path.join(os.path.dirname(gold), name), floating=False, g=0.0)
        assert (m.nb, m.ne, m.ml) == dims
