"""``Trajectory`` yaml files (dynobench/src/motions.cpp:51-132): files stored in the reference tree,
copied VERBATIM to tests/golden/trajectories (make_fixtures.py), must survive read -> write byte
for byte; ``update_feasibility`` reproduces their flags from their figures."""
import glob
import io
import os

import numpy as np
import pytest

from dynoplan_b200.trajectory import Feasibility_thresholds, Trajectory

FILES = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trajectories", "*.yaml")))


def test_fixtures_present():
    assert len(FILES) >= 6


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_round_trip_is_byte_identical(path, tmp_path):
    with open(path) as f:
        raw = f.read()
    t = Trajectory.read_from_yaml(path)
    assert t.to_yaml_format() == raw
    out = tmp_path / "sub" / "traj.yaml"  # create_dir_if_necessary
    t.to_yaml_format(str(out))
    assert out.read_text() == raw
    s = io.StringIO()
    t.to_yaml_format(s, prefix="    ")
    assert s.getvalue() == "".join("    " + ln + "\n" for ln in raw.splitlines())
    assert t.states.shape[0] == int([ln for ln in raw.splitlines() if ln.startswith("num_states")][0].split()[1])


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_update_feasibility_reproduces_the_stored_flags(path):
    t = Trajectory.read_from_yaml(path)
    stored = {k: getattr(t, k) for k in ("feasible", "traj_feas", "goal_feas", "start_feas", "col_feas",
                                         "x_bounds_feas", "u_bounds_feas")}
    if t.max_jump < 0:  # a guess that was never checked: flags are the defaults
        assert not any(stored.values())
        return
    t.update_feasibility(Feasibility_thresholds())
    # figures are printed with 6 digits: a flag may only differ when its figure sits on the threshold
    for k, fig in (("traj_feas", t.max_jump), ("goal_feas", t.goal_distance), ("start_feas", t.start_distance),
                   ("col_feas", t.max_collision), ("x_bounds_feas", t.x_bound_distance), ("u_bounds_feas", t.u_bound_distance)):
        assert getattr(t, k) == stored[k] or abs(fig - 1e-2) < 1e-7, k


def test_result_layouts_and_formatting():
    # read_from_yaml's three accepted layouts (motions.cpp:64-74)
    body = {"states": [[0, 0, 0], [1, 0, .5]], "actions": [[.5, .25]]}
    for node in (body, {"result": body}, {"result": [body]}):
        t = Trajectory.read_from_yaml(node)
        assert t.states.shape == (2, 3) and t.actions.shape == (1, 2)
    with pytest.raises(RuntimeError):
        Trajectory.read_from_yaml({"xs": []})
    # ostream formatting: 6 significant digits, exponents, negative zero, integers without a point
    t = Trajectory(states=np.array([[1e-5, -0.0, 123456789.0, 1 / 3]]), actions=np.zeros((0, 2)), start=np.zeros(2), goal=np.ones(2))
    text = t.to_yaml_format()
    assert "  - [1e-05,-0,1.23457e+08,0.333333]\n" in text
    assert "start: [0,0]\ngoal: [1,1]\n" in text and "cost: 1e+08\n" in text and "num_actions: 0\nactions:\n" in text
