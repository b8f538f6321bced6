"""BASELINE.json's full sizes (configs[2]: 4096 frontier nodes x 100k quadrotor primitives;
configs[3] geometry: 200 boxes) through size-independent properties -- the CSR invariants
nearestR's contract implies, determinism, batch independence -- plus full oracle parity on
a random subset of the nodes (the oracle needs seconds for a few dozen nodes, minutes for
all of them)."""
import numpy as np
import pytest

from dynoplan_b200 import testing
from dynoplan_b200.expander import GpuExpander

pytestmark = pytest.mark.gpu


def _check_invariants(got, case, B):
    lib, robot = case.library, case.robot
    rp = got.row_ptr
    assert rp[0] == 0 and rp[B] == got.n_edges and np.all(np.diff(rp) >= 0)
    # every candidate within the delta bound; rows ascending by (distance, primitive index)
    assert got.delta_dist.max() <= case.radius and got.delta_dist.min() >= 0
    inner = np.ones(got.n_edges, dtype=bool)
    inner[rp[:-1][np.diff(rp) > 0]] = False  # first entry of every non-empty row
    d, p = got.delta_dist, got.prim_idx
    asc = (d[1:] > d[:-1]) | ((d[1:] == d[:-1]) & (p[1:] > p[:-1]))
    assert np.all(asc[inner[1:]])
    assert p.min() >= 0 and p.max() < lib.N
    # no primitive twice in a row's list
    node_of = np.repeat(np.arange(B), np.diff(rp))
    key = node_of.astype(np.int64) * lib.N + p
    assert np.unique(key).shape[0] == key.shape[0]
    # accepted edges own the [V] slots 0..V-1 in CSR order; rejected edges own none
    valid = got.status == 0
    assert got.status.max() <= 3
    assert np.array_equal(got.valid_slot[valid], np.arange(valid.sum()))
    assert np.all(got.valid_slot[~valid] == -1)
    # cost = dt * (steps actually taken); chosen_index inside the primitive
    T = lib.T[p[valid]]
    ch = got.chosen_index
    assert np.all((ch >= -1) & (ch < T))
    steps = np.where(ch >= 0, ch, T - 1)
    assert np.array_equal(got.cost_motion, steps * robot.ref_dt)
    # final states respect the state bounds the reference checks (1e-8 slack, robot_models_base.cpp:417-429)
    lb, ub = np.asarray(robot.x_lb), np.asarray(robot.x_ub)
    assert np.all(got.x_final >= lb - 1e-8) and np.all(got.x_final <= ub + 1e-8)
    assert np.all(got.jump_lbt >= 0) and np.all(got.h >= 0)


@pytest.mark.parametrize("n_prims,n_obs", [(100_000, None), (200_000, 200)])
def test_full_size_expansion_properties(oracle, n_prims, n_obs):
    B = 4096
    case = testing.synthetic_case("quad3d_v0", n_prims, B, seed=0, delta=.8, alpha=.5, n_obs=n_obs)
    rng = np.random.default_rng(7)
    pick = np.sort(rng.choice(B, size=40, replace=False))
    fields = ("row_ptr", "prim_idx", "delta_dist", "status", "valid_slot", "chosen_index", "cost_motion",
              "jump_lbt", "h", "x_final")
    with GpuExpander(case.robot, case.obstacles, case.library, capacity=1 << 24) as ex:
        got = ex.expand(case.nodes, case.radius, case.goal, case.goal_radius, 0)
        again = ex.expand(case.nodes, case.radius, case.goal, case.goal_radius, 0)
        # expand_host returns views of handle-owned pinned memory (valid until the next call): copy
        piped = ex.expand_host(case.nodes, case.radius, case.goal, case.goal_radius, 0).dense()
        piped = type(piped)(*[np.array(getattr(piped, f)) for f in fields])
        sub = ex.expand(case.nodes[pick], case.radius, case.goal, case.goal_radius, 0)
    assert got.n_edges > 100 * B  # a real workload
    _check_invariants(got, case, B)
    # determinism, and the chunk-pipelined host path delivers the same records
    gd = got.dense()
    for name in fields:
        assert np.array_equal(getattr(got, name), getattr(again, name)), name
    for name in fields:
        if name != "valid_slot":
            assert np.array_equal(getattr(gd, name), getattr(piped, name)), name
    # batch independence: a node's list does not depend on what else is in the batch
    sd = sub.dense()
    for k, b in enumerate(pick):
        r_full, r_sub = gd.row(int(b)), sd.row(k)
        assert r_full.stop - r_full.start == r_sub.stop - r_sub.start
        for name in ("prim_idx", "delta_dist", "status", "chosen_index", "cost_motion", "jump_lbt", "h", "x_final"):
            assert np.array_equal(getattr(gd, name)[r_full], getattr(sd, name)[r_sub]), (b, name)
    # oracle parity on the subset
    m = oracle.OracleModel(case.robot, case.obstacles)
    want = oracle.expand(m, oracle.OracleLibrary(case.library), case.nodes[pick], case.radius, case.goal,
                         case.goal_radius, 0, n_threads=8)
    rep = testing.compare_edges(sub, want, case.robot, case.radius)
    print(n_prims, n_obs, rep.summary(), np.bincount(got.status, minlength=4).tolist())
    assert rep.ok, rep.summary()
