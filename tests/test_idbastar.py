"""iDb-A* outer loop (src/idbastar/idbastar.cpp:18-334): schedule, trajectory cutting /
canonicalisation / front insertion on the host; on the GPU every iteration's db-A* must equal
the oracle's search with that iteration's (delta, max_motions), and a library grown at the
front must expand exactly like a fresh handle built from it."""
import os

import numpy as np
import pytest

from dynoplan_b200 import testing
from dynoplan_b200.idbastar import (Options_idbAStar, cut_trajectory, idbA, insert_primitives_front,
                                    make_traj_canonical, schedule)
from dynoplan_b200.motions import load_motion_primitives_new
from dynoplan_b200.robots import robot_factory


def test_schedule_follows_the_reference():
    o = Options_idbAStar(delta_0=.5, num_primitives_0=400, delta_rate=.9, num_primitives_rate=1.5)
    assert schedule(o, []) == (.5, 400)
    # solved -> delta * rate, not solved -> delta * .9999; primitives always * rate (size_t truncation)
    d, n = schedule(o, [True])
    f32 = np.float32  # Options_dbastar::delta is a float: rounded at every update (idbastar.cpp:132-140)
    assert d == float(f32(.5 * .9)) and n == 600
    d, n = schedule(o, [True, False])
    assert d == float(f32(float(f32(.5 * .9)) * .9999)) and n == 900
    assert schedule(Options_idbAStar(num_primitives_0=101, num_primitives_rate=1.5), [False, False])[1] == int(int(101 * 1.5) * 1.5)
    with pytest.raises(NotImplementedError):
        schedule(Options_idbAStar(new_schedule=False), [])


def test_cut_trajectory_matches_the_reference_arithmetic():
    T = 53
    xs, us = np.arange(T * 3, dtype=float).reshape(T, 3), np.arange((T - 1) * 2, dtype=float).reshape(T - 1, 2)
    pieces = cut_trajectory(xs, us, 5)
    length_each = int(T / 5) + 1  # motions.cpp:871
    assert len(pieces) == 5 and all(len(s) == len(a) + 1 for s, a in pieces)
    assert [len(a) for _, a in pieces] == [length_each] * 4 + [T - 1 - 4 * length_each]
    # consecutive pieces share their boundary state; together they cover the trajectory
    for (s0, _), (s1, _) in zip(pieces, pieces[1:]):
        assert np.array_equal(s0[-1], s1[0])
    assert np.array_equal(np.concatenate([a for _, a in pieces]), us)


@pytest.mark.parametrize("system", ["unicycle1_v0", "car1_v0", "quad2d_v0", "quadrotor_v0", "acrobot_v0"])
def test_canonical_pieces_start_at_the_canonical_state(golden_dir, system):
    robot = robot_factory(testing.golden_problem(system, testing.first_env(system)).robotType)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"), robot, 5, add_noise=False)
    s, a = lib.primitive(2)
    # move the primitive somewhere else (the closed-form inverse of the canonicalisation)
    moved = np.array(s)
    ti = robot.translation_invariance
    off = np.linspace(.3, .9, ti)
    moved[:, :ti] += off
    back = make_traj_canonical(robot, moved)
    assert np.allclose(back, make_traj_canonical(robot, s), atol=1e-12)
    assert np.allclose(back[0, :ti], 0)
    if robot.offset_dim > ti:  # quad models: canonical velocity is zero too
        v0 = {"quad2d_v0": slice(3, 5), "quadrotor_v0": slice(7, 10)}[system]
        assert np.allclose(back[0, v0], 0)


@pytest.mark.parametrize("system", ["unicycle1_v0", "unicycle2_v0", "car1_v0", "quad2d_v0", "quadrotor_v0", "acrobot_v0"])
def test_canonical_pieces_are_rollouts_like_the_reference(golden_dir, system):
    """make_trajs_canonical rolls the actions out from canonical_state(x0) (motions.cpp:913-944): a
    piece with dynamics defects (what an optimiser returns) becomes an exact rollout, a defect-free
    one is what the closed form gives."""
    from dynoplan_b200.idbastar import canonical_state
    from oracle import oracle as orc
    robot = robot_factory(testing.golden_problem(system, testing.first_env(system)).robotType)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"), robot, 5, add_noise=False)
    s, a = lib.primitive(1)
    rng = np.random.default_rng(0)
    noisy = s + 1e-3 * rng.normal(size=s.shape)  # defects
    out = make_traj_canonical(robot, noisy, a)
    m = orc.OracleModel(robot)
    x = canonical_state(robot, noisy[0])
    assert np.array_equal(out[0], x)
    for k in range(len(a)):
        x = m.step(x, a[k])
        assert np.allclose(out[k + 1], x, rtol=0, atol=1e-12), k
    # the stored primitives are rollouts themselves (to file precision): same as the closed form
    assert np.allclose(make_traj_canonical(robot, s, a), make_traj_canonical(robot, s), atol=5e-5)


def test_ensure_wraps_angles_and_normalises_quaternions():
    from dynoplan_b200.idbastar import ensure
    x = np.array([1.0, 2.0, 3.5])
    ensure(robot_factory("unicycle1_v0", [0, 0], [6, 6]), x)
    assert abs(x[2] - (3.5 - 2 * np.pi)) < 1e-15
    c = np.array([0.0, 0.0, -4.0, 7.0])
    ensure(robot_factory("car1_v0", [0, 0], [6, 6]), c)
    assert abs(c[2] - (2 * np.pi - 4.0)) < 1e-15 and abs(c[3] - (7.0 - 2 * np.pi)) < 1e-15
    u2 = np.array([0.0, 0.0, 3.5, 0.0, 0.0])
    ensure(robot_factory("unicycle2_v0", [0, 0], [6, 6]), u2)  # the base class's no-op
    assert u2[2] == 3.5
    q = np.zeros(13)
    q[3:7] = [0, 0, .6, .9]
    ensure(robot_factory("quad3d_v0", [-1] * 3, [1] * 3), q)
    assert abs(np.linalg.norm(q[3:7]) - 1) < 1e-15


def test_front_insertion_renumbers(golden_dir):
    robot = robot_factory("unicycle1_v0", [0, 0], [6, 6])
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", "unicycle1_v0.msgpack"), robot, 20, add_noise=False)
    s, a = lib.primitive(7)
    new = insert_primitives_front(lib, robot, cut_trajectory(s, a, 5))
    assert new.N == lib.N + 5 and np.array_equal(new.T[5:], lib.T)
    old_s, old_a = lib.primitive(3)
    got_s, got_a = new.primitive(8)
    assert np.array_equal(old_s, got_s) and np.array_equal(old_a, got_a)
    first_s, first_a = new.primitive(0)
    assert np.allclose(first_s, s[:len(first_s)], atol=2e-7) and not np.array_equal(first_s[0], s[0])  # 1e-7 noise
    assert new.cost[0] == robot.ref_dt * len(first_a)


@pytest.mark.gpu
def test_idbastar_iterations_equal_oracle_searches(oracle, golden_dir):
    from dynoplan_b200.dbastar import Options_dbastar
    from dynoplan_b200.expander import GpuExpander
    p = testing.golden_problem("unicycle1_v0", "bugtrap_0")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", "unicycle1_v0.msgpack"), robot, 300, rand_seed=1)
    oi = Options_idbAStar(delta_0=.3, num_primitives_0=100, max_it=3, timelimit=600)
    od = Options_dbastar(fix_seed=True, search_timelimit=120e3)
    with GpuExpander(robot, p.obstacles, lib) as ex:
        info = idbA(ex, lib, p.start, p.goal, oi, od, trajectory_optimization=None)
    assert info.exit_criteria == "max_it" and len(info.trajs_raw) == 3 and not info.solved
    m, ol = oracle.OracleModel(robot, p.obstacles), oracle.OracleLibrary(lib)
    hist = []
    for k, out in enumerate(info.trajs_raw):
        delta, n = schedule(oi, hist)
        assert info.schedule[k] == (delta, n)
        want = oracle.dbastar(m, ol, p.start, p.goal, delta=delta, max_motions=n, search_timelimit=120e3)
        assert out.solved == bool(want.summary["solved"])
        assert int(out.data["expands"]) == want.summary["expands"]
        assert np.array_equal(out.motions_used, want.motions_used)
        assert abs(out.cost - want.summary["cost"]) < 1e-4
        hist.append(out.solved)
    assert info.solved_raw == any(hist)


@pytest.mark.gpu
def test_idbastar_feeds_solutions_back_into_the_gpu_store(oracle, golden_dir):
    from dynoplan_b200.dbastar import Options_dbastar
    from dynoplan_b200.expander import GpuExpander
    p = testing.golden_problem("unicycle1_v0", "parallelpark_0")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", "unicycle1_v0.msgpack"), robot, 300, rand_seed=1)
    calls = []

    def fake_optimizer(xs, us):  # stands in for crocoddyl: accepts the db-A* plan as it is
        calls.append(len(us))
        return xs, us, True, robot.ref_dt * len(us)
    oi = Options_idbAStar(delta_0=.3, num_primitives_0=100, max_it=2, max_num_sol=5, timelimit=600)
    with GpuExpander(robot, p.obstacles, lib) as ex:
        info = idbA(ex, lib, p.start, p.goal, oi, Options_dbastar(fix_seed=True, search_timelimit=120e3), fake_optimizer)
        assert info.solved and len(calls) >= 1 and info.library_size == lib.N + 5 * len(calls)
        # the handle now holds the grown library: expansion == a fresh handle built from it
        grown_first = ex.get_traj(np.zeros(robot.nx), 0)
        nodes = testing.frontier_nodes(robot, lib, 64, p.p_lb, p.p_ub, seed=3)
        got = ex.expand(nodes, .15, p.goal, .3, 0).dense()
    with GpuExpander(robot, p.obstacles, info.library) as fresh:
        want = fresh.expand(nodes, .15, p.goal, .3, 0).dense()
        assert np.array_equal(grown_first, fresh.get_traj(np.zeros(robot.nx), 0))
    assert info.states is not None and abs(info.cost - robot.ref_dt * len(info.actions)) < 1e-9
    assert got.n_edges == want.n_edges > 0
    for name in ("row_ptr", "prim_idx", "delta_dist", "status", "chosen_index", "cost_motion", "jump_lbt", "h", "x_final"):
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
    # and the oracle agrees on the grown library
    rep = testing.compare_edges(got, oracle.expand(oracle.OracleModel(robot, p.obstacles), oracle.OracleLibrary(info.library),
                                                   nodes, .15, p.goal, .3, 0), robot)
    assert rep.ok, rep.summary()
