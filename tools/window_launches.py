"""One solve of the quadrotor window instance (first 3000 shipped primitives, delta .8): for an ncu launch list
of what a db-A* batch costs per kernel at benchmark-instance size (tools/launch_sums.py reads it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynoplan_b200 import testing
from dynoplan_b200.dbastar import Options_dbastar, dbastar
from dynoplan_b200.expander import GpuExpander
from dynoplan_b200.robots import robot_factory
p = testing.golden_problem("quadrotor_v0", "window")
robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
lib = testing.load_compact_library(os.path.join(testing.GOLDEN, "motions", "quadrotor_v0_3000_compact.npz"), robot, 3000, rand_seed=1)
with GpuExpander(robot, p.obstacles, lib) as ex:
    r = dbastar(ex, lib, p.start, p.goal, Options_dbastar(fix_seed=True, delta=.8, max_motions=3000, limit_branching_factor=30,
                                                          search_timelimit=600e3, tn_device=int(os.environ.get("TN", -1))))
    print(r.solved, r.data["expands"], r.data["gpu_calls"], r.data["time_gpu"], r.time_search)
