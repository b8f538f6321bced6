#!/usr/bin/env python
"""BASELINE.json configs[4]: the Dynobench sweep -- every problem of the six systems, independent
problems distributed over the GPUs with NO communication (one worker process per GPU, problem i goes
to worker i mod N; the reference distributes them over CPU processes the same way,
benchmark/benchmark.py:1395-1445, 1630-1675).

    python tools/sweep.py [--gpus N] [--out sweep.json] [--check]

Per problem: the db-A* stages of iDb-A* with the per-system parameters of
benchmark/config/algs/idbastar_v0.yaml (delta_0, num_primitives_0, limit_branching_factor,
delta_factor_goal, search_timelimit) and its schedule (not solved -> delta * .9999, primitives * 1.5,
idbastar.cpp:125-146) until the first solved search; trajectory optimisation (crocoddyl) is out of
scope, so the reported cost is the raw db-A* plan's.  Libraries: the shipped
dynomotions/*.small5000.msgpack when the reference checkout is present, else the committed subsets
(tests/golden/motions: 300-400 primitives, the quadrotor's first 3000 as a compact fixture).
--check: every search is repeated by the CPU oracle and must be identical."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_MOTIONS = "/root/reference/dynomotions"

# benchmark/config/algs/idbastar_v0.yaml:10-139 (default search_timelimit 15 s, limit_branching_factor 20)
PARAMS = {
    "unicycle1_v0": dict(delta_0=.3, num_primitives_0=100),
    "unicycle2_v0": dict(delta_0=.5, num_primitives_0=200),
    "car1_v0": dict(delta_0=.3, num_primitives_0=400),
    "quad2d_v0": dict(delta_0=.5, num_primitives_0=400, limit_branching_factor=20, search_timelimit=45e3),
    "acrobot_v0": dict(delta_0=.35, num_primitives_0=100000, limit_branching_factor=40, delta_factor_goal=.7,
                       search_timelimit=25e3),
    "quadrotor_v0": dict(delta_0=.7, num_primitives_0=2000, limit_branching_factor=30, search_timelimit=60e3),
}
SHIPPED = {
    "unicycle1_v0": "unicycle1_v0__ispso__2023_04_03__14_56_57.bin.im.bin.im.bin.small5000.msgpack",
    "unicycle2_v0": "unicycle2_v0__ispso__2023_04_03__15_36_01.bin.im.bin.im.bin.small5000.msgpack",
    "car1_v0": "car1_v0_all.bin.sp.bin.small5000.msgpack",
    "quad2d_v0": "quad2d_v0_all_im.bin.sp.bin.ca.bin.small5000.msgpack",
    "acrobot_v0": "acrobot_v0_all2.bin.sp.bin.small5000.msgpack",
    "quadrotor_v0": "quad3d_v0_all3.bin.im.bin.sp1.bin.ca.bin.small5000.msgpack",
}


def problems():
    out = []
    for system in sorted(PARAMS):
        d = os.path.join(GOLDEN, "envs", system)
        for f in sorted(os.listdir(d)):
            if f.endswith(".yaml"):
                with open(os.path.join(d, f)) as fh:
                    if "\nrobots:" in fh.read():  # car1_v0/empty.yaml is an environment without a problem
                        out.append((system, f[:-5]))
    return out


def load_library(system, robot):
    from dynoplan_b200 import testing
    from dynoplan_b200.motions import load_motion_primitives_new
    shipped = os.path.join(REF_MOTIONS, SHIPPED[system])
    if os.path.exists(shipped):
        return load_motion_primitives_new(shipped, robot, rand_seed=1), "shipped small5000"
    if system == "quadrotor_v0":
        return testing.load_compact_library(os.path.join(GOLDEN, "motions", "quadrotor_v0_3000_compact.npz"), robot,
                                            rand_seed=1), "committed compact fixture (first 3000)"
    return load_motion_primitives_new(os.path.join(GOLDEN, "motions", system + ".msgpack"), robot, rand_seed=1), \
        "committed subset"


def worker(rank: int, world: int, check: bool, max_it: int):
    from dynoplan_b200 import testing
    from dynoplan_b200.dbastar import Options_dbastar, dbastar
    from dynoplan_b200.expander import GpuExpander
    from dynoplan_b200.robots import robot_factory
    libs = {}
    for i, (system, env) in enumerate(problems()):
        if i % world != rank:
            continue
        p = testing.golden_problem(system, env)
        robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
        if system not in libs:
            libs[system] = load_library(system, robot)
        lib, lib_src = libs[system]
        par = PARAMS[system]
        rec = {"system": system, "problem": env, "gpu": rank, "library": f"{lib.N} primitives, {lib_src}", "iterations": []}
        t0 = time.perf_counter()
        with GpuExpander(robot, p.obstacles, lib, device=0) as ex:
            delta, n = np.float32(par["delta_0"]), int(par["num_primitives_0"])
            for it in range(max_it):
                n_use = min(n, lib.N)
                opt = Options_dbastar(fix_seed=True, delta=float(delta), max_motions=n_use,
                                      limit_branching_factor=par.get("limit_branching_factor", 20),
                                      delta_factor_goal=par.get("delta_factor_goal", 1.0),
                                      search_timelimit=par.get("search_timelimit", 15e3))
                out = dbastar(ex, lib, p.start, p.goal, opt)
                row = {"delta": float(delta), "primitives": n_use, "solved": bool(out.solved), "cost": float(out.cost),
                       "expands": int(out.data["expands"]), "time_search_ms": float(out.time_search),
                       "time_gpu_ms": float(out.data["time_gpu"]), "status": out.data["terminate_status"],
                       "feasible": out.data["feasible"] == "1"}
                if check:
                    from oracle import oracle as orc
                    want = orc.dbastar(orc.OracleModel(robot, p.obstacles), orc.OracleLibrary(lib), p.start, p.goal,
                                       delta=float(delta), max_motions=n_use,
                                       limit_branching_factor=opt.limit_branching_factor,
                                       delta_factor_goal=opt.delta_factor_goal, search_timelimit=600e3,
                                       max_expands=int(out.data["expands"]) if not out.solved else 1000000)
                    row["identical_to_oracle"] = bool(int(out.data["expands"]) == want.summary["expands"]
                                                      and np.array_equal(out.motions_used, want.motions_used))
                    row["oracle_time_search_ms"] = want.summary["time_search_ms"]
                rec["iterations"].append(row)
                if out.solved or n_use >= lib.N and it > 0:
                    break
                delta = np.float32(float(delta) * .9999)  # not solved (idbastar.cpp:136-140)
                n = int(n * 1.5)
        last = rec["iterations"][-1]
        rec.update(solved=last["solved"], cost=last["cost"] if last["solved"] else None,
                   expands=sum(r["expands"] for r in rec["iterations"]),
                   time_search_ms=sum(r["time_search_ms"] for r in rec["iterations"]),
                   wall_s=time.perf_counter() - t0)
        print("SWEEP " + json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--max-it", type=int, default=6)
    ap.add_argument("--worker", type=int, default=None)
    a = ap.parse_args()
    if a.worker is not None:
        worker(a.worker, a.gpus, a.check, a.max_it)
        return
    t0 = time.perf_counter()
    procs = []
    for r in range(a.gpus):
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--gpus", str(a.gpus), "--worker", str(r),
                                       "--max-it", str(a.max_it)] + (["--check"] if a.check else []),
                                      stdout=subprocess.PIPE, env=env, text=True))
    recs = []
    for pr in procs:
        out, _ = pr.communicate()
        recs += [json.loads(line[6:]) for line in out.splitlines() if line.startswith("SWEEP ")]
        if pr.returncode != 0:
            print("a worker failed with exit code", pr.returncode, file=sys.stderr)
    wall = time.perf_counter() - t0
    recs.sort(key=lambda r: (r["system"], r["problem"]))
    expands = sum(r["expands"] for r in recs)
    search_s = sum(r["time_search_ms"] for r in recs) / 1e3
    summary = {"gpus": a.gpus, "problems": len(recs), "solved": sum(r["solved"] for r in recs), "expands": expands,
               "sum_time_search_s": search_s, "expansions_per_s_in_search": expands / max(search_s, 1e-9),
               "wall_s": wall, "problems_per_s": len(recs) / wall,
               "all_identical_to_oracle": (all(it.get("identical_to_oracle", False) for r in recs for it in r["iterations"])
                                           if a.check else None)}
    for r in recs:
        print(f"{r['system']:13s} {r['problem']:24s} gpu {r['gpu']} solved {int(r['solved'])} cost {r['cost']} "
              f"its {len(r['iterations'])} expands {r['expands']:7d} search {r['time_search_ms']:9.1f} ms "
              f"({r['expands'] / max(r['time_search_ms'], 1e-9) * 1e3:8.0f} exp/s)"
              + (f" identical {all(it['identical_to_oracle'] for it in r['iterations'])}" if a.check else ""))
    print(json.dumps(summary))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            json.dump({"summary": summary, "problems": recs}, f, indent=1)
    sys.exit(0 if all(pr.returncode == 0 for pr in procs) else 1)


if __name__ == "__main__":
    main()
