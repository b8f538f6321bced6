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


def test_options_read_from_yaml_and_print(tmp_path):
    """Options_dbastar / Options_idbAStar: read_from_yaml with the keys at the top level or under
    options_dbastar / options_idbastar (options.cpp:55-73, idbastar.hpp:87-109), unknown keys
    ignored, absent keys keep their value; print() in the reference's key order and formatting."""
    from dynoplan_b200.dbastar import Options_dbastar
    from dynoplan_b200.idbastar import Options_idbAStar
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text("options_dbastar:\n  delta: .25\n  max_motions: 400\n  fix_seed: 1\n  heuristic: 1\n  solver_id: 1\n"
                   "options_idbastar:\n  delta_0: .5\n  num_primitives_0: 200\n  new_schedule: false\n")
    o = Options_dbastar().read_from_yaml(str(cfg))
    assert (o.delta, o.max_motions, o.fix_seed, o.heuristic, o.alpha) == (.25, 400, True, 1, .5)
    assert isinstance(o.max_motions, int) and isinstance(o.fix_seed, bool)
    i = Options_idbAStar().read_from_yaml(str(cfg))
    assert (i.delta_0, i.num_primitives_0, i.new_schedule, i.max_it) == (.5, 200, False, 10)
    flat = Options_dbastar().read_from_yaml({"delta": .7, "search_timelimit": 15000, "smooth_traj": True})
    assert flat.delta == .7 and flat.search_timelimit == 15000.0
    s = io.StringIO()
    i.print(s, "  ")
    assert s.getvalue() == ("  delta_0: 0.5\n  num_primitives_0: 200\n  delta_rate: 0.9\n  num_primitives_rate: 1.5\n"
                            "  timelimit: 10\n  max_it: 10\n  new_schedule: 0\n  max_motions_primitives: 10000\n"
                            "  add_primitives_opt: 1\n")
    s = io.StringIO()
    o.print(s)
    lines = s.getvalue().splitlines()
    assert lines[0] == "fix_seed: 1" and lines[4] == "delta: 0.25" and lines[-1] == "check_cols: 1" and len(lines) == 26
