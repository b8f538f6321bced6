"""Oracle of the roadmap heuristic (db-A* heuristic = 1; SURVEY 8(f) N2 -- not yet on the
product path) against the one test the reference holds for it:
test/dbastar/test_dbastar_0.cpp:64-116 (test_heu_map) -- on unicycle1 bugtrap_0 the values
h(goal), h(5,5,0), h(4,5.5,1.57), h(1,5,0), h(start) must already be in ascending order."""
import numpy as np
import pytest

from dynoplan_b200 import testing
from dynoplan_b200.robots import robot_factory


def test_heu_map_orders_the_reference_probe_states(oracle):
    p = testing.golden_problem("unicycle1_v0", "bugtrap_0")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    m = oracle.OracleModel(robot, p.obstacles)
    # options of the reference test: max_size_heu_map = 2000, num_sample_trials = 4000
    samples = oracle.HeuristicMap.sample(m, robot, p.start, p.goal, num_sample_trials=4000, max_size_heu_map=2000, seed=0)
    assert np.array_equal(samples[-1], np.asarray(p.goal, dtype=float)) and len(samples) > 1000
    hm = oracle.HeuristicMap(m, samples)  # heu_connection_radius 1, heu_resolution .5 (options.hpp:120,141)
    assert hm.d[-1] == 0 and hm.n_edges > len(samples)
    probes = [p.goal, [5, 5, 0], [4, 5.5, 1.57], [1, 5, 0], p.start]
    h = [hm.h(x, p.goal, connect_radius_h=.5) for x in probes]
    assert h == sorted(h), h
    assert h[0] < 1e-9 and h[-1] < 1e8  # the goal costs nothing; the start is connected
    # the roadmap bound dominates the euclidean one behind the bugtrap's walls
    assert h[-1] > m.lower_bound_time(p.start, p.goal)


def test_roadmap_building_blocks(oracle):
    p = testing.golden_problem("unicycle1_v0", "bugtrap_0")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    m = oracle.OracleModel(robot, p.obstacles)
    L = oracle.lib()
    a, b = np.array([.5, .5, 3.0]), np.array([.5, .5, -3.0])
    x = np.zeros(3)
    L.orc_interpolate(m._h, oracle._dp(x), oracle._dp(a), oracle._dp(b), .5)  # so2 interpolation wraps through pi
    assert abs(abs(x[2]) - np.pi) < .15 and x[0] == .5
    # an edge through a wall is rejected, a free one accepted
    free = [o for o in p.obstacles]
    c = np.array(free[0].center[:2])
    a, b = np.r_[c + [-1.5, 0], 0.0], np.r_[c + [1.5, 0], 0.0]
    if L.orc_collision_check(m._h, oracle._dp(a)) and L.orc_collision_check(m._h, oracle._dp(b)):
        assert not L.orc_check_edge_at_resolution(m._h, oracle._dp(a), oracle._dp(b), .5)
    assert L.orc_check_edge_at_resolution(m._h, oracle._dp(np.asarray(p.start, float)), oracle._dp(np.asarray(p.start, float)), .5)
    # acrobot: the reference has no lower_bound_time_pr (NOT_IMPLEMENTED); unicycle2 zeroes v, w
    p2 = testing.golden_problem("unicycle2_v0", "parallelpark_0")
    r2 = robot_factory(p2.robotType, p2.p_lb, p2.p_ub)
    m2 = oracle.OracleModel(r2, p2.obstacles)
    s2 = oracle.HeuristicMap.sample(m2, r2, p2.start, p2.goal, num_sample_trials=300, max_size_heu_map=200, seed=1)
    assert np.all(s2[:, 3:] == 0)
    hm2 = oracle.HeuristicMap(m2, s2)
    moving = np.array(p2.start, dtype=float)
    moving[3] = .4
    assert hm2.h(moving, p2.goal) >= hm2.h(p2.start, p2.goal)  # velocity bound can only raise h


def test_product_dijkstra_and_file_format_against_the_oracle_map(oracle, tmp_path):
    """Host side of N2 in the product (dynoplan_b200/heuristics.py): Dijkstra over the roadmap's
    edge list and the reference's heu-map file format, against the oracle's map."""
    from dynoplan_b200 import heuristics
    p = testing.golden_problem("unicycle1_v0", "bugtrap_0")
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    m = oracle.OracleModel(robot, p.obstacles)
    samples = oracle.HeuristicMap.sample(m, robot, p.start, p.goal, num_sample_trials=600, max_size_heu_map=250, seed=2)
    hm = oracle.HeuristicMap(m, samples)
    L = oracle.lib()
    edges, weights = [], []
    for i in range(len(samples)):
        for j in range(i + 1, len(samples)):
            if m.distance(samples[i], samples[j]) < 1.0 and \
                    L.orc_check_edge_at_resolution(m._h, oracle._dp(samples[i]), oracle._dp(samples[j]), .5):
                edges.append((i, j))
                weights.append(m.lower_bound_time(samples[i], samples[j]))
    assert len(edges) == hm.n_edges
    d, par = heuristics.compute_heuristic_map(len(samples), edges, weights)
    assert np.allclose(d, hm.d, rtol=1e-12, atol=0) and d[-1] == 0
    reach = d < heuristics.UNREACHABLE
    assert np.array_equal(par[~reach], np.arange(len(samples))[~reach])
    # following the parents reaches the goal with exactly the stored cost
    w = {e: c for e, c in zip(edges, weights)}
    for v in np.flatnonzero(reach)[:40]:
        cost, u = 0.0, int(v)
        while u != len(samples) - 1:
            nxt = int(par[u])
            cost += w[(min(u, nxt), max(u, nxt))]
            u = nxt
        assert abs(cost - d[v]) < 1e-9
    path = str(tmp_path / "heu.yaml")
    heuristics.write_heu_map(path, samples, d, par, header="# test")
    xs2, d2, p2 = heuristics.load_heu_map(path)
    assert np.array_equal(xs2, samples) and np.array_equal(d2, d) and np.array_equal(p2, par)
    with pytest.raises(RuntimeError):
        (tmp_path / "bad.yaml").write_text("other: 1\n")
        heuristics.load_heu_map(str(tmp_path / "bad.yaml"))


@pytest.mark.parametrize("system,env", [("car1_v0", "parallelpark_0"), ("quad2d_v0", "quad_obs_column"),
                                        ("quadrotor_v0", "recovery_with_obs")])
def test_roadmap_oracle_on_the_other_models(oracle, system, env):
    """Sanity of the oracle's roadmap for the models the reference test does not cover: the goal
    costs nothing, h never drops below the velocity bound, connected probes cost at least the
    straight-line bound."""
    p = testing.golden_problem(system, env)
    robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
    m = oracle.OracleModel(robot, p.obstacles)
    samples = oracle.HeuristicMap.sample(m, robot, p.start, p.goal, num_sample_trials=500, max_size_heu_map=300, seed=3)
    if robot.kind == 5:
        assert np.allclose(np.linalg.norm(samples[1:-1, 3:7], axis=1), 1)
    hm = oracle.HeuristicMap(m, samples, heu_connection_radius=2.0, heu_resolution=.5)
    R = 1.5
    goal0 = samples[-1]
    assert hm.h(goal0, p.goal, R) <= hm.h(p.start, p.goal, R)
    assert hm.h(goal0, goal0, R) < 1e-9
    for x in samples[1:40]:
        hx = hm.h(x, p.goal, R)
        if hx < 1e8:
            assert hx >= m.lower_bound_time(np.r_[x], np.r_[goal0]) - 1e-9  # triangle inequality over the roadmap
