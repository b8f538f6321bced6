"""Primitive file formats: the readers against the file
PAIRS shipped in the reference tree -- the same libraries stored as .bin (boost binary
archive) and .yaml / .msgpack (tests/golden/formats, copied unchanged) -- and the writers
against the readers."""
import os

import numpy as np
import pytest

from dynoplan_b200 import motions
from dynoplan_b200.boost_archive import read_trajectories_boost
from dynoplan_b200.motions import MotionLibrary, load_motion_primitives_new, save_motion_primitives
from dynoplan_b200.robots import robot_factory

HERE = os.path.dirname(os.path.abspath(__file__))
FORMATS = os.path.join(HERE, "golden", "formats")
YAML_PAIR = "motions__unicycle1_v0__20-07-2023--16-04-22"
MSGPACK_PAIR = "unicycle1_v0__ispso__2023_04_03__14_56_57.bin.less.bin"


def test_bin_equals_its_yaml_twin():
    a = read_trajectories_boost(os.path.join(FORMATS, YAML_PAIR + ".bin"))
    b = motions._read_trajs(os.path.join(FORMATS, YAML_PAIR + ".yaml"))
    assert len(a) == len(b) == 10
    for (s, u, c), (s2, u2, c2) in zip(a, b):
        s2, u2 = np.array(s2, dtype=float), np.array(u2, dtype=float)
        assert s.shape == s2.shape and u.shape == u2.shape and s.shape[0] == u.shape[0] + 1
        # the yaml twin was printed with 6 significant digits
        assert np.allclose(s, s2, rtol=1e-5, atol=1e-6) and np.allclose(u, u2, rtol=1e-5, atol=1e-6)
        # only the boost archive carries a cost; the yaml reader ignores a stored one
        # (Trajectory::read_from_yaml, motions.cpp:51-75): the trajectory keeps its default
        assert c2 == 1e8 and c < 1e5


def test_bin_equals_its_msgpack_twin_bit_for_bit():
    a = read_trajectories_boost(os.path.join(FORMATS, MSGPACK_PAIR))
    b = motions._read_trajs(os.path.join(FORMATS, MSGPACK_PAIR + ".msgpack"))
    assert len(a) == len(b) == 100
    for (s, u, _), (s2, u2, _) in zip(a, b):
        assert np.array_equal(s, np.array(s2, dtype=float)) and np.array_equal(u, np.array(u2, dtype=float))


def test_loader_goes_through_every_format():
    robot = robot_factory("unicycle1_v0", [0, 0], [6, 6])
    lb = load_motion_primitives_new(os.path.join(FORMATS, MSGPACK_PAIR), robot, 40, rand_seed=1)
    lm = load_motion_primitives_new(os.path.join(FORMATS, MSGPACK_PAIR + ".msgpack"), robot, 40, rand_seed=1)
    assert lb.N == lm.N == 40
    for name in ("T", "states", "actions", "cost"):
        assert np.array_equal(getattr(lb, name), getattr(lm, name)), name
    with pytest.raises(RuntimeError):
        load_motion_primitives_new(os.path.join(FORMATS, "nope.txt"), robot)


@pytest.mark.parametrize("ext", [".msgpack", ".json", ".bin", ".yaml"])
def test_writers_round_trip(tmp_path, ext):
    robot = robot_factory("unicycle1_v0", [0, 0], [6, 6])
    lib = load_motion_primitives_new(os.path.join(FORMATS, MSGPACK_PAIR + ".msgpack"), robot, 25, add_noise=False)
    out = str(tmp_path / ("lib" + ext))
    save_motion_primitives(out, lib)
    back = load_motion_primitives_new(out, robot, add_noise=False)
    assert back.N == lib.N and np.array_equal(back.T, lib.T)
    assert np.array_equal(back.states, lib.states) and np.array_equal(back.actions, lib.actions)
    assert np.allclose(back.cost, lib.cost)
    # truncated write
    save_motion_primitives(out, lib, num_motions=7)
    assert load_motion_primitives_new(out, robot, add_noise=False).N == 7


def test_bin_writer_reproduces_the_reference_file():
    src = os.path.join(FORMATS, MSGPACK_PAIR)
    trajs = read_trajectories_boost(src)
    lib = MotionLibrary.from_lists([(s, u) for s, u, _ in trajs], 3, 2, cost=[c for _, _, c in trajs])
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "x.bin")
        save_motion_primitives(out, lib)
        mine, ref = open(out, "rb").read(), open(src, "rb").read()
    # identical up to the archive's trailing byte
    assert mine == ref[:len(mine)] and len(ref) - len(mine) <= 1
