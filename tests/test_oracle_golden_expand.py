"""The oracle against its own committed expansion records (regression pin; the same
fixtures are the golden vectors of the CUDA path in test_gpu_parity.py)."""
import os

import numpy as np
import pytest

from dynoplan_b200 import testing
from dynoplan_b200.motions import load_motion_primitives_new
from dynoplan_b200.robots import robot_factory

SYSTEMS = ["unicycle1_v0", "unicycle2_v0", "car1_v0", "quad2d_v0", "acrobot_v0", "quadrotor_v0"]


def load_case(golden_dir, system):
    g = np.load(os.path.join(golden_dir, f"expand_{system}.npz"))
    p = testing.golden_problem(system, str(g["env"]))
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"), robot, rand_seed=1)
    return g, p, robot, lib


@pytest.mark.parametrize("system", SYSTEMS)
def test_oracle_matches_golden(oracle, golden_dir, system):
    g, p, robot, lib = load_case(golden_dir, system)
    m = oracle.OracleModel(robot, p.obstacles)
    out = oracle.expand(m, oracle.OracleLibrary(lib), g["nodes"], float(g["radius"]), g["goal"],
                        float(g["goal_radius"]), 0)
    assert np.array_equal(out.row_ptr, g["row_ptr"])
    assert np.array_equal(out.prim_idx, g["prim_idx"])
    assert np.array_equal(out.delta_dist, g["delta_dist"])
    assert np.array_equal(out.status, g["status"])
    ok = out.status == 0
    assert np.array_equal(out.chosen_index[ok], g["chosen_index"][ok])
    assert np.allclose(out.x_final[ok], g["x_final"][ok], rtol=0, atol=1e-13)
    assert np.allclose(out.h[ok], g["h"][ok], rtol=0, atol=1e-13)
    # candidates are sorted by ascending delta distance inside every row (a5)
    for b in range(len(out.row_ptr) - 1):
        d = out.delta_dist[out.row_ptr[b]:out.row_ptr[b + 1]]
        assert np.all(np.diff(d) >= 0) and np.all(d <= float(g["radius"]))


def test_threads_give_identical_results(oracle, golden_dir):
    g, p, robot, lib = load_case(golden_dir, "quad2d_v0")
    m = oracle.OracleModel(robot, p.obstacles)
    ol = oracle.OracleLibrary(lib)
    a = oracle.expand(m, ol, g["nodes"], float(g["radius"]), g["goal"], float(g["goal_radius"]), 0)
    b = oracle.expand(m, ol, g["nodes"], float(g["radius"]), g["goal"], float(g["goal_radius"]), 0,
                      n_threads=4)
    assert np.array_equal(a.prim_idx, b.prim_idx) and np.array_equal(a.status, b.status)
    assert np.array_equal(a.x_final, b.x_final)


def test_rollout_equals_stored_states(oracle, golden_dir):
    # stored quad3d / quad2d states are exact Euler rollouts of the stored
    # actions with the model's step, and primitives start at p = v = 0
    for system in ("quadrotor_v0", "quad2d_v0"):
        g, p, robot, lib = load_case(golden_dir, system)
        raw = load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"),
                                         robot, add_noise=False)
        m = oracle.OracleModel(robot, [])
        worst = 0.0
        for i in range(0, raw.N, 17):
            xs, us = raw.primitive(i)
            x = xs[0].copy()
            for k in range(len(us)):
                x = m.step(x, us[k])
                worst = max(worst, float(np.max(np.abs(x - xs[k + 1]))))
        assert worst < 5e-7, (system, worst)  # float32-encoded entries limit this
