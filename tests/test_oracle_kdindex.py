"""The CPU baseline's sub-linear delta query (oracle/kdindex.cpp: a k-d index with per-subtree lower
bounds of the nigh metric, standing in for the reference's nigh tree,
include/dynoplan/nigh_custom_spaces.hpp:236-247) returns EXACTLY what the linear scan returns --
same candidates, same distances, same order -- on all six models, radii from tiny to all-inclusive,
node states outside the angle range, libraries with duplicate keys."""
import os

import numpy as np
import pytest

from dynoplan_b200 import testing
from dynoplan_b200.motions import load_motion_primitives_new
from dynoplan_b200.robots import robot_factory


@pytest.mark.parametrize("system", ["unicycle1_v0", "unicycle2_v0", "car1_v0", "quad2d_v0", "acrobot_v0", "quadrotor_v0"])
def test_kd_filter_equals_linear_scan(oracle, golden_dir, system):
    p = testing.golden_problem(system, testing.first_env(system))
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"), robot, rand_seed=1)
    m, ol = oracle.OracleModel(robot, p.obstacles), oracle.OracleLibrary(lib)
    nodes = testing.frontier_nodes(robot, lib, 96, p.p_lb, p.p_ub, seed=2)
    nodes[:5] += 40 * np.pi * np.array([1.0 if k in (2,) and robot.nx <= 6 else 0.0 for k in range(robot.nx)])  # angles far outside [-pi, pi)
    total = 0
    for radius in (1e-3, .05, .2, .4, 1.5, 50.0):
        for n_use in (lib.N, 37):
            a = oracle.expand(m, ol, nodes, radius, p.goal, .3, 0, n_use=n_use, filter_only=True)
            b = oracle.expand(m, ol, nodes, radius, p.goal, .3, 0, n_use=n_use, filter_only=True, kd_filter=True, n_threads=3)
            assert np.array_equal(a.row_ptr, b.row_ptr) and np.array_equal(a.prim_idx, b.prim_idx), (radius, n_use)
            assert np.array_equal(a.delta_dist, b.delta_dist)
            total += int(a.row_ptr[-1])
    assert total > 1000


def test_kd_filter_on_a_synthetic_quadrotor_library_and_duplicates(oracle):
    robot = robot_factory("quad3d_v0", [-5] * 3, [5] * 3)
    lib = oracle.generate_library(robot, 6000, 11, 40, seed=4)
    q = lib.states[:, 3:7]
    q /= np.sqrt((q * q).sum(axis=1, keepdims=True))
    off = lib.state_off
    lib.states[off[100:140]] = lib.states[off[100]]  # forty identical keys: a subtree of zero extent
    m, ol = oracle.OracleModel(robot, []), oracle.OracleLibrary(lib)
    nodes = testing.frontier_nodes(robot, lib, 64, [-5] * 3, [5] * 3, seed=9)
    nodes[0] = lib.states[off[100]]
    for radius in (.1, .4, .8):
        a = oracle.expand(m, ol, nodes, radius, nodes[0], .8, 0, filter_only=True, fast_filter=True)
        b = oracle.expand(m, ol, nodes, radius, nodes[0], .8, 0, filter_only=True, kd_filter=True)
        assert np.array_equal(a.row_ptr, b.row_ptr) and np.array_equal(a.prim_idx, b.prim_idx)
        assert np.array_equal(a.delta_dist, b.delta_dist)
    assert a.row_ptr[1] - a.row_ptr[0] >= 40
