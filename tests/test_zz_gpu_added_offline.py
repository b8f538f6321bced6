"""GPU cases written while no GPU was available (end of round 1): they run VALIDATED kernels on new
data, so they are expected to pass, but they have not run yet -- the file name sorts last so that
under ``pytest -x`` a surprise here cannot hide the validated suites."""
import os

import numpy as np
import pytest

from dynoplan_b200 import testing
from dynoplan_b200.motions import load_motion_primitives_new, reverse_motions
from dynoplan_b200.robots import robot_factory


@pytest.mark.gpu
@pytest.mark.parametrize("system,env", [("unicycle1_v0", "bugtrap_0"), ("unicycle2_v0", "parallelpark_0")])
def test_gpu_reverse_expander_unicycles(oracle, golden_dir, system, env):
    """Backward expansion for the unicycles = the forward kernels over the reversed library
    (LazyTraj::compute(forward = false) is the plain transform_primitive for them)."""
    from dynoplan_b200.expander import GpuExpander
    p = testing.golden_problem(system, env)
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    lib = load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"), robot, 300, rand_seed=1)
    rev = reverse_motions(lib, robot, rand_seed=2)
    nodes = testing.frontier_nodes(robot, rev, 128, p.p_lb, p.p_ub, seed=7)
    radius = .25
    with GpuExpander(robot, p.obstacles, rev) as ex:
        got = ex.expand(nodes, radius, p.start, .3, heuristic=0)  # a backward tree grows towards the START
    want = oracle.expand(oracle.OracleModel(robot, p.obstacles), oracle.OracleLibrary(rev), nodes, radius, p.start, .3, 0)
    rep = testing.compare_edges(got, want, robot, radius)
    assert rep.ok, rep.summary()
    assert got.n_edges > 100
