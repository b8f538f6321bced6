"""Option structs of the callers added in round 2: defaults, yaml keys and print order follow the reference
(include/dynoplan/dbrrt/dbrrt.hpp:11-107, include/dynoplan/tdbastar/options.hpp:18-60, src/tdbastar/options.cpp)."""
import io

from dynoplan_b200.dbrrt import Options_dbrrt
from dynoplan_b200.tdbastar import Constraint, Options_tdbastar, constraints_from_yaml


def test_options_dbrrt_defaults_yaml_and_print():
    o = Options_dbrrt()
    # dbrrt.hpp:12-40
    assert (o.use_connect, o.prob_expand_forward, o.seed, o.cost_jump, o.max_step_size) == (False, .5, -1, 1.0, 3.0)
    assert (o.goal_region, o.goal_bias, o.delta, o.max_expands, o.timelimit, o.max_motions) == (.3, .1, .3, 10000, 10000.0, 1000)
    assert o.choose_first_motion_valid and not o.do_optimization and not o.ao_rrt
    o.read_from_yaml({"options_dbastar": {"delta": .5, "seed": 3, "use_connect": 1, "goal_bias": .9}})
    assert (o.delta, o.seed, o.use_connect) == (.5, 3, True)
    assert o.goal_bias == .1  # not among the keys Options_dbrrt::__load_data sets (dbrrt.hpp:50-87)
    out = io.StringIO()
    o.print(out)
    lines = out.getvalue().splitlines()
    assert lines[0] == "do_optimization: 0" and lines[1] == "cost_jump: 1" and lines[-1] == "use_collision_shape: 0"
    assert len(lines) == len(Options_dbrrt.yaml_fields) == 26


def test_options_tdbastar_and_constraints(golden_dir):
    o = Options_tdbastar()
    assert o.rewire is True and o.delta == .3 and o.limit_branching_factor == 20  # options.hpp:19-57
    o.read_from_yaml({"options_tdbastar": {"delta": .75, "max_motions": 400, "rewire": False}})
    assert (o.delta, o.max_motions) == (.75, 400) and o.rewire is True  # rewire is not a yaml key (options.cpp:16-41)
    cons = constraints_from_yaml(golden_dir + "/constraints/car1_v0.yaml")
    assert len(cons) == 12 and isinstance(cons[0], Constraint)
    assert abs(cons[0].time - 1.8) < 1e-6 and cons[0].time != 1.8  # read through float, like `.as<float>()`
    assert cons[3].time == 2.0 and len(cons[0].constrained_state) == 4
