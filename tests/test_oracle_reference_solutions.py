"""Pins the oracle's whole db-A* loop against OUTPUTS OF THE REFERENCE ITSELF: the db-A*
solutions stored in the reference tree (dynobench/envs/**/idbastar_v0_db_solution_v0.yaml),
produced by dynoplan with the first ``num_primitives_0`` primitives and ``delta_0`` of
benchmark/config/algs/idbastar_v0.yaml.  The oracle reproduces them state by state
(6 printed digits).  tests/golden/make_fixtures.py copied them here."""
import os

import numpy as np
import pytest
import yaml

from dynoplan_b200 import testing
from dynoplan_b200.motions import load_motion_primitives_new
from dynoplan_b200.robots import robot_factory

# system -> (num_primitives_0, delta_0, extra options)   [idbastar_v0.yaml:10-139]
CFG = {"unicycle1_v0": (100, .3, {}), "unicycle2_v0": (200, .5, {}), "car1_v0": (400, .3, {}),
       "quad2d_v0": (400, .5, {"limit_branching_factor": 20})}
CASES = [("unicycle1_v0", "bugtrap_0"), ("unicycle1_v0", "parallelpark_0"),
         ("unicycle2_v0", "kink_0"), ("unicycle2_v0", "parallelpark_0"),
         ("car1_v0", "kink_0"), ("car1_v0", "parallelpark_0"),
         ("quad2d_v0", "empty_0"), ("quad2d_v0", "empty_1"), ("quad2d_v0", "quad_obs_column")]


@pytest.mark.parametrize("system,env", CASES)
def test_oracle_reproduces_stored_reference_solution(oracle, golden_dir, system, env):
    n_prims, delta, extra = CFG[system]
    p = testing.golden_problem(system, env)
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    # rand_seed: the loader's 1e-7 noise comes from the process-global rand() stream, which other
    # tests advance -- a fresh reference process starts at srand(1)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"),
                                     robot, n_prims, rand_seed=1)
    m = oracle.OracleModel(robot, p.obstacles)
    out = oracle.dbastar(m, oracle.OracleLibrary(lib), p.start, p.goal, delta=delta,
                         max_motions=n_prims, search_timelimit=120e3, **extra)
    with open(os.path.join(golden_dir, "ref_solutions", f"{system}__{env}.yaml")) as f:
        ref = yaml.safe_load(f)
    s = out.summary
    assert s["solved"] == 1 and s["feasible"] == 1 and s["err"] == 0
    assert abs(s["cost"] - ref["cost"]) < 1e-9
    assert s["n_states"] == ref["num_states"]
    want = np.array(ref["states"], dtype=float)
    # the reference prints 6 significant digits (math_utils.hpp:10, FMT)
    assert np.max(np.abs(out.states - want)) < 1e-5 * (1 + np.abs(want).max())
    assert np.max(np.abs(out.actions - np.array(ref["actions"], dtype=float))) < 1e-5
    for key in ("max_jump", "goal_distance", "start_distance"):
        # 6 significant digits in the file (half a unit of the last one) + the loader noise along the plan
        assert abs(s[key] - ref[key]) < 6e-6, key
    assert s["max_collision"] == ref["max_collision"] == 0


def test_reference_test_bugtrap_heu(oracle, golden_dir):
    # test/dbastar/test_dbastar_0.cpp:118-163: unicycle1 bugtrap, max_motions=100,
    # default delta, heuristic 0 -> solved and cost < 100
    p = testing.golden_problem("unicycle1_v0", "bugtrap_0")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", "unicycle1_v0.msgpack"), robot, 100)
    m = oracle.OracleModel(robot, p.obstacles)
    for heu in (0, -1):
        out = oracle.dbastar(m, oracle.OracleLibrary(lib), p.start, p.goal, max_motions=100,
                             heuristic=heu, search_timelimit=1e5)
        assert out.summary["solved"] == 1 and out.summary["cost"] < 100


def test_options_validation(oracle, golden_dir):
    # dbastar.cpp:588-594: alpha must be in (0, 1)
    p = testing.golden_problem("unicycle1_v0", "parallelpark_0")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", "unicycle1_v0.msgpack"), robot, 50)
    m = oracle.OracleModel(robot, p.obstacles)
    out = oracle.dbastar(m, oracle.OracleLibrary(lib), p.start, p.goal, alpha=1.0)
    assert out.summary["err"] == 2
    # max_expands -> MAX_EXPANDS (status 1), unsolved: trajectory to the closest node
    out = oracle.dbastar(m, oracle.OracleLibrary(lib), p.start, p.goal, max_expands=3)
    assert out.summary["status"] == 1 and out.summary["solved"] == 0 and out.summary["n_states"] >= 1
