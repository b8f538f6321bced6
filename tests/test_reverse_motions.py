"""``reverse_motions`` (src/dbrrt/dbrrt.cpp:252-340): the library a backward search expands from.
CPU: structure and exactness against a restatement with the oracle's ``step``; the reversed
library walked FORWARD from its (canonical) first state with the reversed actions does not
reproduce the states (time runs backwards) -- what must hold is that every reversed primitive is
the original one, translated, in reverse order.  The GPU case (an expander over the reversed
unicycle library is the reverse expander) lives in tests/test_zz_gpu_added_offline.py."""
import os

import numpy as np
import pytest

from dynoplan_b200 import testing
from dynoplan_b200.motions import load_motion_primitives_new, reverse_motions
from dynoplan_b200.robots import robot_factory


def _lib(golden_dir, system, n=40):
    p = testing.golden_problem(system, testing.first_env(system))
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    return p, robot, load_motion_primitives_new(os.path.join(golden_dir, "motions", system + ".msgpack"), robot, n, rand_seed=1)


@pytest.mark.parametrize("system", ["unicycle1_v0", "unicycle2_v0"])
def test_reverse_unicycles_is_an_exact_translation(golden_dir, system):
    _, robot, lib = _lib(golden_dir, system)
    rev = reverse_motions(lib, robot, add_noise=False)
    assert np.array_equal(rev.T, lib.T) and rev.N == lib.N and np.all(rev.cost == 1e8)
    for i in range(lib.N):
        xs, us = lib.primitive(i)
        rx, ru = rev.primitive(i)
        want = xs.copy()
        want[:, :2] -= xs[-1, :2]
        assert np.array_equal(rx, want[::-1]) and np.array_equal(ru, us[::-1])
        assert np.all(rx[0, :2] == 0)  # the reversed primitive starts at the canonical position
    # with the loader's noise: first and last state only, 1e-7 * Random from libc rand()
    noisy = reverse_motions(lib, robot, rand_seed=3)
    d = noisy.states - rev.states
    off = lib.state_off
    ends = np.zeros(lib.states.shape[0], dtype=bool)
    ends[off[:-1]] = ends[off[1:] - 1] = True
    assert np.all(d[~ends] == 0) and np.all(np.abs(d[ends]) <= 1e-7) and np.abs(d[ends]).max() > 1e-9
    assert np.array_equal(noisy.states, reverse_motions(lib, robot, rand_seed=3).states)  # deterministic per seed


@pytest.mark.parametrize("system", ["quad2d_v0", "quadrotor_v0"])
def test_reverse_rotor_models_re_rollout(oracle, golden_dir, system):
    _, robot, lib = _lib(golden_dir, system, 12)
    rev = reverse_motions(lib, robot, add_noise=False)
    m = oracle.OracleModel(robot, [])
    ti, v0 = robot.translation_invariance, (7 if robot.kind == 5 else 3)
    for i in range(lib.N):
        xs, us = lib.primitive(i)
        T = xs.shape[0]
        vel = xs[-1, v0:v0 + ti]
        x = xs[0].copy()
        x[:ti] -= xs[-1, :ti] - (T - 1) * robot.ref_dt * vel
        x[v0:v0 + ti] -= vel
        want = [x]
        for k in range(T - 1):
            want.append(m.step(want[-1], us[k]))
        want = np.array(want)[::-1]
        rx, ru = rev.primitive(i)
        if robot.kind == 5:  # all quaternions are normalised at the end (dbrrt.cpp:316-321)
            want[:, 3:7] /= np.sqrt((want[:, 3:7] ** 2).sum(axis=1, keepdims=True))
        assert np.array_equal(rx, want) and np.array_equal(ru, us[::-1])
        # the reversed primitive starts at rest at the origin (up to the rollout's rounding)
        assert np.abs(rx[0, :ti]).max() < 1e-5 and np.abs(rx[0, v0:v0 + ti]).max() < 1e-5
        # and is the original one, shifted by the closed form p - p_off - k*dt*v_T, v - v_T
        k = np.arange(T)[:, None]
        closed = xs.copy()
        closed[:, :ti] += -(xs[-1, :ti] - (T - 1) * robot.ref_dt * vel) - k * robot.ref_dt * vel
        closed[:, v0:v0 + ti] -= vel
        assert np.abs(rx[::-1][:, :ti] - closed[:, :ti]).max() < 2e-5  # file precision of the stored body


def test_reverse_rejects_other_models(golden_dir):
    for system in ("car1_v0", "acrobot_v0"):
        _, robot, lib = _lib(golden_dir, system, 5)
        with pytest.raises(RuntimeError):
            reverse_motions(lib, robot)
