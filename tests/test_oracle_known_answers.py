"""Pins the CPU oracle against every known-answer vector the reference's own tests hold
for this path (dynobench/test/test_models.cpp) -- SURVEY.md section 8(c)."""
import math

import numpy as np
import pytest

from dynoplan_b200.problem import Problem
from dynoplan_b200.robots import robot_from_params, robot_factory
from dynoplan_b200 import testing


def _model(oracle, params, system, env):
    p = testing.golden_problem(system, env)
    return oracle.OracleModel(robot_from_params(params), p.obstacles)


def test_tcol_unicycle(oracle):
    # dynobench/test/test_models.cpp:773-803 (Model_unicycle1() defaults, parallelpark_0)
    m = _model(oracle, {"dynamics": "unicycle1"}, "unicycle1_v0", "parallelpark_0")
    assert abs(m.collision_distance([.7, .8, 0]) - .25) < 1e-7
    assert abs(m.collision_distance([1.9, .3, 0]) - .3) < 1e-7
    assert abs(m.collision_distance([1.5, .3, .1]) - (-0.11123)) < 1e-5


def test_tcol_quad3d(oracle):
    # dynobench/test/test_models.cpp:833-868 (Model_quad3d() default sphere radius 0.4)
    m = _model(oracle, {"dynamics": "quad3d"}, "quadrotor_v0", "quad_one_obs")
    tail = [0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    assert abs(m.collision_distance([1., 1., 1.] + tail) - 0.824745) < 1e-5
    assert abs(m.collision_distance([1.2, 1.5, 2.] + tail) - (-0.0999999)) < 1e-5
    assert abs(m.collision_distance([5., 5., 1.] + tail) - 0.824745) < 1e-5
    assert abs(m.collision_distance([3., 1., 3.] + tail) - .1) < 5e-5


def test_col_acrobot(oracle):
    # dynobench/test/test_models.cpp:870-904
    m = _model(oracle, {"dynamics": "acrobot"}, "acrobot_v0", "swing_up_obs")
    tol = 1e-4
    assert abs(m.collision_distance([0, 0, 0, 0]) - 1.59138) < tol
    assert abs(m.collision_distance([3.14159, 0, 0, 0]) - 1.1) < tol
    assert abs(m.collision_distance([math.pi / 2, math.pi / 2, 0, 0]) - 1.59138) < tol
    assert abs(m.collision_distance([2.37, 1.4, 0, 0]) - 1.21897) < tol


def test_collision_check_sign_agrees_with_distance(oracle):
    # boolean contact (collide) and signed distance must agree away from the boundary
    rng = np.random.default_rng(0)
    for system, env in [("unicycle1_v0", "bugtrap_0"), ("car1_v0", "kink_0"),
                        ("quadrotor_v0", "window"), ("acrobot_v0", "swing_up_obs")]:
        p = testing.golden_problem(system, env)
        robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
        m = oracle.OracleModel(robot, p.obstacles)
        lo = np.array(p.p_lb + [-3.0] * (robot.nx - len(p.p_lb)))
        hi = np.array(p.p_ub + [3.0] * (robot.nx - len(p.p_ub)))
        n_free = 0
        for _ in range(300):
            x = rng.uniform(lo, hi)
            if robot.kind == 5:
                x[3:7] /= np.linalg.norm(x[3:7])
            d = m.collision_distance(x)
            if abs(d) > 1e-9:
                assert m.collision_check(x) == (d > 0)
                n_free += d > 0
        assert n_free > 0


def test_quat_product_matches_hamilton(oracle):
    # t_quat_product (dynobench/test/test_models.cpp:452-527): quat_product == Eigen's
    # product.  The quad3d step composes q * Exp(w dt); check it through step().
    robot = robot_factory("quad3d_v0")
    m = oracle.OracleModel(robot, [])
    rng = np.random.default_rng(1)

    def hamilton(p, q):  # (x, y, z, w)
        px, py, pz, pw = p
        qx, qy, qz, qw = q
        return np.array([pw * qx + px * qw + py * qz - pz * qy,
                         pw * qy - px * qz + py * qw + pz * qx,
                         pw * qz + px * qy - py * qx + pz * qw,
                         pw * qw - px * qx - py * qy - pz * qz])

    for _ in range(50):
        x = np.zeros(13)
        q = rng.normal(size=4)
        x[3:7] = q / np.linalg.norm(q)
        x[10:13] = rng.uniform(-8, 8, size=3)
        u = rng.uniform(0, 1.3, size=4)
        xn = m.step(x, u)
        th = np.linalg.norm(x[10:13] * robot.ref_dt)
        dq = np.concatenate([x[10:13] * robot.ref_dt / th * math.sin(th / 2), [math.cos(th / 2)]])
        assert np.allclose(xn[3:7], hamilton(x[3:7], dq), atol=1e-10)
        assert abs(np.linalg.norm(xn[3:7]) - 1) < 1e-9
        # exp map small-angle branch (exp_map_quat, test_models.cpp:529-619)
    x = np.zeros(13)
    x[6] = 1
    x[10:13] = [1e-6, -2e-6, 1e-6]
    xn = m.step(x, np.ones(4))
    assert np.allclose(xn[3:7], [5e-9, -1e-8, 5e-9, 1.0], atol=1e-12)


def test_hover_is_equilibrium(oracle):
    # u = 1 is hover for the rotor models (u_nominal = m g / n_rotors)
    for name in ("quad3d_v0", "quad2d_v0"):
        robot = robot_factory(name)
        m = oracle.OracleModel(robot, [])
        x = np.zeros(robot.nx)
        if name == "quad3d_v0":
            x[6] = 1
        xn = m.step(x, np.ones(robot.nu))
        assert np.allclose(xn, x, atol=1e-12)


def test_bisection_indices(oracle):
    # is_motion_collision_free (dynobench/src/motions.cpp:1042-1106)
    for dt, res in ((.01, 2), (.1, 1)):
        for T in range(2, 130):
            idx = oracle.bisection_indices(T, dt)
            assert idx[0] == 0 and idx[1] == T - 1
            s = sorted(set(idx))
            assert len(s) == len(idx) or T <= 2  # every index visited once
            gaps = np.diff(s)
            assert gaps.max() <= res, (T, dt, s)
    assert oracle.bisection_indices(5, .1) == [0, 4, 2, 3, 1]
    assert oracle.bisection_indices(8, .01) == [0, 7, 3, 5, 1]


def test_nigh_metric_properties(oracle):
    rng = np.random.default_rng(2)
    for name in ("unicycle1_v0", "unicycle2_v0", "car1_v0", "quad2d_v0", "acrobot_v0", "quad3d_v0"):
        robot = robot_factory(name)
        m = oracle.OracleModel(robot, [])
        for _ in range(20):
            a, b = rng.uniform(-3, 3, size=(2, robot.nx))
            if robot.kind == 5:
                a[3:7] /= np.linalg.norm(a[3:7])
                b[3:7] /= np.linalg.norm(b[3:7])
            # acos(|q.q|) of a not-exactly-unit product: nigh returns ~sqrt(2 eps), as the reference does
            assert m.nigh_distance(a, a) <= (1e-7 if robot.kind == 5 else 0)
            assert m.nigh_distance(a, b) == m.nigh_distance(b, a) >= 0
    # unicycle1: 1 * |dp| + .5 * so2, so2 wraps (nigh/impl/so2.hpp:47-51)
    m = oracle.OracleModel(robot_factory("unicycle1_v0"), [])
    assert abs(m.nigh_distance([0, 0, 3.0], [3, 4, -3.0]) - (5 + .5 * (2 * math.pi - 6))) < 1e-12
    # quad3d: q and -q are the same rotation
    m = oracle.OracleModel(robot_factory("quad3d_v0"), [])
    a = np.zeros(13)
    a[6] = 1
    b = a.copy()
    b[6] = -1
    assert m.nigh_distance(a, b) == 0


def test_canonical_and_offset(oracle):
    x = np.arange(1, 14, dtype=float)
    m = oracle.OracleModel(robot_factory("quad3d_v0"), [])
    assert np.array_equal(m.offset(x), [1, 2, 3, 8, 9, 10])
    c = m.canonical_state(x)
    assert np.array_equal(c[[0, 1, 2, 7, 8, 9]], np.zeros(6)) and np.array_equal(c[3:7], x[3:7])
    m = oracle.OracleModel(robot_factory("quad2d_v0"), [])
    assert np.array_equal(m.offset(x[:6]), [1, 2, 4, 5])
    m = oracle.OracleModel(robot_factory("acrobot_v0"), [])
    assert np.array_equal(m.canonical_state(x[:4]), x[:4]) and len(m.offset(x[:4])) == 0
    m = oracle.OracleModel(robot_factory("car1_v0"), [])
    assert np.array_equal(m.canonical_state(x[:4]), [0, 0, 3, 4])
