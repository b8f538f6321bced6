"""db-A* on the quadrotor BENCHMARK INSTANCES, solved to completion (BASELINE.json north_star:
"identical db-A* solutions on the unicycle/car/quadrotor benchmark instances"; the reference's own
cases are test/dbastar/test_dbastar_0.cpp:166-280: quad_one_obs and window with the first 3000
primitives of quad3d_v0_all3...small5000.msgpack at delta = .8, and
benchmark/config/algs/idbastar_v0.yaml:109-139).

Library: tests/golden/motions/quadrotor_v0_3000_compact.npz, the first 3000 primitives of that
file without their interior states (make_fixtures.py explains why nothing reads them); where the
reference checkout exists the compact library is checked against the file itself."""
import os

import numpy as np
import pytest

from dynoplan_b200 import testing
from dynoplan_b200.robots import robot_factory

REF_FILE = "/root/reference/dynomotions/quad3d_v0_all3.bin.im.bin.sp1.bin.ca.bin.small5000.msgpack"
# instance -> (primitives, expansions / plan cost of the oracle's search, recorded when the fixture was made)
INSTANCES = {"window": (3000, 8270, 2.54), "recovery": (3000, 7711, 3.46), "recovery_with_obs": (3000, 8912, 3.72),
             "quad_one_obs": (2000, 31566, 3.93), "empty_0_easy": (3000, 118, 1.6)}
KW = dict(delta=.8, limit_branching_factor=30, search_timelimit=600e3)


_CACHE = {}


def _library(golden_dir, robot, n):
    """The first n primitives (loaded once per session: the noise of primitive i does not depend on
    how many follow it, so the head of the 3000 is the library a load of n gives)."""
    if "lib" not in _CACHE:
        _CACHE["lib"] = testing.load_compact_library(os.path.join(golden_dir, "motions", "quadrotor_v0_3000_compact.npz"),
                                                     robot, 3000, rand_seed=1)
    return _CACHE["lib"].head(n)


def test_compact_library_is_the_reference_file(oracle, golden_dir):
    """First / last states, actions and lengths are the file's; the searches are the ones the file
    gives (interior states are re-rolled out by the reference, never read from the file)."""
    if not os.path.exists(REF_FILE):
        pytest.skip("the reference checkout is not on this machine")
    from dynoplan_b200.motions import load_motion_primitives_new
    p = testing.golden_problem("quadrotor_v0", "window")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    full = load_motion_primitives_new(REF_FILE, robot, 3000, rand_seed=1)
    comp = _library(golden_dir, robot, 3000)
    off = full.state_off
    assert np.array_equal(full.T, comp.T) and np.array_equal(full.actions, comp.actions)
    assert np.array_equal(full.states[off[:-1]], comp.states[off[:-1]])
    assert np.array_equal(full.states[off[1:] - 1], comp.states[off[1:] - 1])
    assert np.max(np.abs(full.states - comp.states)) < 1e-12  # the file's interiors ARE rollouts
    m = oracle.OracleModel(robot, p.obstacles)
    a = oracle.dbastar(m, oracle.OracleLibrary(full), p.start, p.goal, max_motions=3000, **KW)
    b = oracle.dbastar(m, oracle.OracleLibrary(comp), p.start, p.goal, max_motions=3000, **KW)
    assert a.summary["solved"] and a.summary["expands"] == b.summary["expands"] == INSTANCES["window"][1]
    assert np.array_equal(a.motions_used, b.motions_used) and np.array_equal(a.states, b.states)


@pytest.mark.parametrize("env", ["empty_0_easy", "window"])
def test_oracle_solves_the_instances(oracle, golden_dir, env):
    n, expands, cost = INSTANCES[env]
    p = testing.golden_problem("quadrotor_v0", env)
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = _library(golden_dir, robot, n)
    r = oracle.dbastar(oracle.OracleModel(robot, p.obstacles), oracle.OracleLibrary(lib), p.start, p.goal,
                       max_motions=n, **KW)
    assert r.summary["solved"] and r.summary["expands"] == expands and abs(r.summary["cost"] - cost) < 1e-9
    assert r.summary["feasible"]


@pytest.mark.gpu
@pytest.mark.parametrize("env", list(INSTANCES))
def test_gpu_search_equals_the_oracle_on_the_quadrotor_instances(oracle, golden_dir, env):
    """Same expansions, same primitive sequence, same plan (states to 1e-9, cost to 1e-4 -- the
    north star's tolerance; observed: equal), solved and feasible."""
    from dynoplan_b200.dbastar import Options_dbastar, dbastar
    from dynoplan_b200.expander import GpuExpander
    n, expands, cost = INSTANCES[env]
    p = testing.golden_problem("quadrotor_v0", env)
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = _library(golden_dir, robot, n)
    with GpuExpander(robot, p.obstacles, lib) as ex:
        got = dbastar(ex, lib, p.start, p.goal, Options_dbastar(fix_seed=True, max_motions=n, **KW))
    want = oracle.dbastar(oracle.OracleModel(robot, p.obstacles), oracle.OracleLibrary(lib), p.start, p.goal,
                          max_motions=n, **KW)
    s = want.summary
    assert s["solved"] and s["expands"] == expands
    assert got.solved and int(got.data["expands"]) == s["expands"] and int(got.data["all_nodes"]) == s["n_nodes"]
    assert np.array_equal(got.motions_used, want.motions_used)
    assert abs(got.cost - s["cost"]) < 1e-4 and abs(got.cost - cost) < 1e-4
    assert got.states.shape == want.states.shape and np.allclose(got.states, want.states, rtol=0, atol=1e-9)
    assert got.data["feasible"] == "1"
    print(env, "expands", got.data["expands"], "gpu calls", got.data["gpu_calls"],
          "search %.0f ms (oracle %.0f ms)" % (got.time_search, s["time_search_ms"]))
