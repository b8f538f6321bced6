"""Roadmap heuristic on the GPU path (db-A* heuristic = 1, SURVEY 8(f) N2, query side):
``Heu_roadmap::h`` of every accepted edge and of arbitrary states against the oracle
(oracle/heuristic.cpp, pinned to the reference's test_heu_map).

First ran on a B200 at the start of round 2 (profiles/r2_pytest_heuristic_first_gpu_run.log: 4 passed);
the per-thread device functions are also tested, compiled for the host, in
tests/test_device_source_on_host.py."""

import numpy as np
import pytest

from dynoplan_b200 import _abi, testing
from dynoplan_b200.expander import GpuExpander
from dynoplan_b200.robots import robot_factory

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("system,env", [("unicycle1_v0", "bugtrap_0"), ("unicycle2_v0", "parallelpark_0"),
                                        ("quad2d_v0", "quad_obs_column")])
def test_roadmap_h_matches_oracle(oracle, golden_dir, system, env):
    import os
    from dynoplan_b200.motions import load_motion_primitives_new
    p = testing.golden_problem(system, env)
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"), robot, 200, rand_seed=1)
    m = oracle.OracleModel(robot, p.obstacles)
    samples = oracle.HeuristicMap.sample(m, robot, p.start, p.goal, num_sample_trials=1500, max_size_heu_map=600, seed=0)
    hm = oracle.HeuristicMap(m, samples)
    d = np.where(hm.d > 1e300, 1e300, hm.d)  # unreachable nodes keep boost's "infinite" distance
    nodes = testing.frontier_nodes(robot, lib, 96, p.p_lb, p.p_ub, seed=4)
    with GpuExpander(robot, p.obstacles, lib) as ex:
        with pytest.raises(_abi.DynoError):  # no map yet
            ex.expand(nodes, .2, p.goal, .3, heuristic=1)
        ex.set_heuristic_map(samples, d, connect_radius_h=.5)
        # arbitrary states
        got = ex.heuristic(nodes, p.goal)
        want = np.array([hm.h(x, p.goal, .5) for x in nodes])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12), np.abs(got - want).max()
        # per accepted edge of an expansion
        e = ex.expand(nodes, .2, p.goal, .3, heuristic=1)
        assert e.n_valid > 0
        want_e = np.array([hm.h(x, p.goal, .5) for x in e.x_final])
        assert np.allclose(e.h, want_e, rtol=1e-12, atol=1e-12)
        # everything else is what heuristic = 0 delivers
        e0 = ex.expand(nodes, .2, p.goal, .3, heuristic=0)
        for name in ("row_ptr", "prim_idx", "status", "chosen_index", "x_final", "jump_lbt"):
            assert np.array_equal(getattr(e, name), getattr(e0, name)), name
        ex.set_heuristic_map(np.zeros((0, robot.nx)), np.zeros(0))
        with pytest.raises(_abi.DynoError):
            ex.expand(nodes, .2, p.goal, .3, heuristic=1)


def test_roadmap_build_matches_oracle(oracle, golden_dir):
    """Map build on the GPU path: sample filter, O(n^2) edge test, Dijkstra -- against the oracle's
    map over the SAME samples; then a whole db-A* solve with heuristic = 1."""
    import os
    from dynoplan_b200 import heuristics
    from dynoplan_b200.dbastar import Options_dbastar, dbastar
    from dynoplan_b200.motions import load_motion_primitives_new
    p = testing.golden_problem("unicycle1_v0", "bugtrap_0")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", "unicycle1_v0.msgpack"), robot, 100, rand_seed=1)
    m = oracle.OracleModel(robot, p.obstacles)
    with GpuExpander(robot, p.obstacles, lib) as ex:
        rng = np.random.default_rng(5)
        probes = np.array([heuristics.sample_uniform(robot, rng) for _ in range(2000)])
        assert np.array_equal(ex.collision_check(probes), np.array([m.collision_check(x) for x in probes]))
        xs, d, par = heuristics.generate_heuristic_map(ex, robot, p.start, p.goal, num_sample_trials=1200,
                                                       max_size_heu_map=500, seed=0)
        hm = oracle.HeuristicMap(m, xs)
        assert np.allclose(d, hm.d, rtol=1e-12, atol=0) and d[-1] == 0
        ex.set_heuristic_map(xs, np.minimum(d, 1e300), connect_radius_h=.5)
        got = dbastar(ex, lib, p.start, p.goal, Options_dbastar(fix_seed=True, delta=.3, max_motions=100, heuristic=1,
                                                                search_timelimit=120e3))
        base = dbastar(ex, lib, p.start, p.goal, Options_dbastar(fix_seed=True, delta=.3, max_motions=100, heuristic=0,
                                                                 search_timelimit=120e3))
    assert got.solved and base.solved
    # the roadmap heuristic is the better informed one behind the bugtrap's walls (on the CPU:
    # 2563 against 13148 expansions, tests/test_host_loop.py::test_host_loop_with_the_roadmap_heuristic)
    assert int(got.data["expands"]) < int(base.data["expands"])
    # ... and the search is the oracle's with Heu_roadmap over the same map
    hm.attach(.5)
    want = oracle.dbastar(m, oracle.OracleLibrary(lib), p.start, p.goal, delta=.3, max_motions=100, heuristic=1,
                          search_timelimit=120e3)
    assert int(got.data["expands"]) == want.summary["expands"]
    assert np.array_equal(got.motions_used, want.motions_used)
