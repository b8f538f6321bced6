"""Regenerates tests/golden/ from the reference checkout (run in the authoring
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_fixtures.py

Produces
  envs/<system>/<problem>.yaml      problem files (data) of the six benchmark systems
  motions/<system>.msgpack          the first K primitives of the shipped
                                    dynomotions/*.small5000.msgpack libraries, re-encoded
                                    like nlohmann::json::to_msgpack (float32 when lossless)
  formats/*                         the reference's format pairs for loader parity
                                    (.bin <-> .yaml, .bin <-> .msgpack)
  ref_solutions/<system>__<problem>.yaml
                                    db-A* solutions stored in the reference tree
                                    (dynobench/envs/**/idbastar_v0_db_solution_v0.yaml):
                                    outputs of the REFERENCE ITSELF, used to pin the oracle
  expand_<system>.npz               oracle expansion records for a fixed node batch
  motions/quadrotor_v0_3000_compact.npz
                                    the first 3000 quadrotor primitives (what test/dbastar/
                                    test_dbastar_0.cpp:224-229 searches with) WITHOUT their interior
                                    states: T, first state, last state, actions.  The reference
                                    re-rolls a quadrotor primitive out from its (transformed) first
                                    state (transform_primitive2, robot_models_base.cpp:656-689) and
                                    reads the stored last state only in the closed-form pre-check, so
                                    a library whose interior states are that rollout gives the same
                                    searches as the file (tests/test_quadrotor_instances.py checks it
                                    against the file where the reference checkout exists); the file's
                                    13 x 167k full-precision doubles would be 15 MB.
"""
import glob
import os
import shutil
import struct
import sys

import msgpack
import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/"

SYSTEMS = {  # system dir -> (library file, primitives kept)
    "unicycle1_v0": ("unicycle1_v0__ispso__2023_04_03__14_56_57.bin.im.bin.im.bin.small5000.msgpack", 300),
    "unicycle2_v0": ("unicycle2_v0__ispso__2023_04_03__15_36_01.bin.im.bin.im.bin.small5000.msgpack", 400),
    "car1_v0": ("car1_v0_all.bin.sp.bin.small5000.msgpack", 400),
    "quad2d_v0": ("quad2d_v0_all_im.bin.sp.bin.ca.bin.small5000.msgpack", 400),
    "acrobot_v0": ("acrobot_v0_all2.bin.sp.bin.small5000.msgpack", 300),
    "quadrotor_v0": ("quad3d_v0_all3.bin.im.bin.sp1.bin.ca.bin.small5000.msgpack", 400),
}

# stored reference db-A* solutions the oracle reproduces state by state with the first
# num_primitives_0 primitives and delta_0 of benchmark/config/algs/idbastar_v0.yaml
REF_SOLUTIONS = [
    ("unicycle1_v0", "bugtrap_0"), ("unicycle1_v0", "parallelpark_0"),
    ("unicycle2_v0", "kink_0"), ("unicycle2_v0", "parallelpark_0"),
    ("car1_v0", "kink_0"), ("car1_v0", "parallelpark_0"),
    ("quad2d_v0", "empty_0"), ("quad2d_v0", "empty_1"), ("quad2d_v0", "quad_obs_column"),
]


def pack_float(v: float) -> bytes:
    f32 = struct.unpack(">f", struct.pack(">f", v))[0] if abs(v) < 3e38 else None
    if f32 is not None and f32 == v:
        return b"\xca" + struct.pack(">f", v)
    return b"\xcb" + struct.pack(">d", v)


def pack(obj) -> bytes:
    if isinstance(obj, float):
        return pack_float(obj)
    if isinstance(obj, dict):
        out = msgpack.packb({})[:0]
        n = len(obj)
        out += bytes([0x80 | n]) if n < 16 else b"\xde" + struct.pack(">H", n)
        for k, v in obj.items():
            out += msgpack.packb(k) + pack(v)
        return out
    if isinstance(obj, (list, tuple)):
        n = len(obj)
        if n < 16:
            head = bytes([0x90 | n])
        elif n < 65536:
            head = b"\xdc" + struct.pack(">H", n)
        else:
            head = b"\xdd" + struct.pack(">I", n)
        return head + b"".join(pack(v) for v in obj)
    return msgpack.packb(obj)


def main():
    for d in ("envs", "motions", "formats", "ref_solutions"):
        os.makedirs(os.path.join(HERE, d), exist_ok=True)
    # problems
    for system in SYSTEMS:
        os.makedirs(os.path.join(HERE, "envs", system), exist_ok=True)
        for f in sorted(glob.glob(REF + f"dynobench/envs/{system}/*.yaml")):
            shutil.copy(f, os.path.join(HERE, "envs", system, os.path.basename(f)))
    # truncated libraries
    for system, (fname, keep) in SYSTEMS.items():
        with open(REF + "dynomotions/" + fname, "rb") as f:
            data = msgpack.unpackb(f.read(), raw=False)["data"][:keep]
        blob = pack({"data": [{"actions": t["actions"], "states": t["states"]} for t in data]})
        back = msgpack.unpackb(blob, raw=False)["data"]
        assert all(a["states"] == b["states"] and a["actions"] == b["actions"] for a, b in zip(data, back))
        with open(os.path.join(HERE, "motions", system + ".msgpack"), "wb") as f:
            f.write(blob)
        print(system, keep, "primitives", len(blob) // 1024, "KiB")
    # format pairs
    for f in (glob.glob(REF + "data/motions__unicycle1_v0__*") +
              glob.glob(REF + "data/motion_primitives/unicycle1_v0/*")):
        shutil.copy(f, os.path.join(HERE, "formats", os.path.basename(f)))
    # stored reference solutions (states + summary numbers only)
    for system, env in REF_SOLUTIONS:
        src = REF + f"dynobench/envs/{system}/{env}/idbastar_v0_db_solution_v0.yaml"
        with open(src) as f:
            sol = yaml.load(f, Loader=yaml.CSafeLoader)
        keep = {k: sol[k] for k in ("cost", "max_jump", "max_collision", "goal_distance",
                                    "start_distance", "num_states", "states", "actions")}
        with open(os.path.join(HERE, "ref_solutions", f"{system}__{env}.yaml"), "w") as f:
            yaml.dump(keep, f, Dumper=yaml.CSafeDumper, default_flow_style=None, width=200)
    # a few stored trajectory files verbatim: the byte-level pin of Trajectory::to_yaml_format
    os.makedirs(os.path.join(HERE, "trajectories"), exist_ok=True)
    for rel in RAW_TRAJECTORIES:
        shutil.copy(REF + "dynobench/envs/" + rel, os.path.join(HERE, "trajectories", rel.replace("/", "__")))
    # the constraint files of the reference's tdb-A* tests (test/tdbastar/test_tdbastar_0.cpp:150-231), verbatim;
    # their problems (envs/*/swap/swap1_*.yaml: empty [0, 5]^2, start (1, 2.5, 0..), goal (4, 2.5, 0..)) are
    # spelled out in the tests
    os.makedirs(os.path.join(HERE, "constraints"), exist_ok=True)
    for src, dst in (("unicycle1_v0/constraints/constraints_unicycle.yaml", "unicycle1_v0.yaml"),
                     ("unicycle2_v0/constraints/constraints_unicycle2.yaml", "unicycle2_v0.yaml"),
                     ("car1_v0/constraints/constraints_trailer.yaml", "car1_v0.yaml")):
        shutil.copy(REF + "dynobench/envs/" + src, os.path.join(HERE, "constraints", dst))
    make_expand_fixtures()


RAW_TRAJECTORIES = ["unicycle1_v0/parallelpark_0/idbastar_v0_db_solution_v0.yaml",
                    "unicycle1_v0/parallelpark_0/idbastar_v0_solution_v0.yaml",   # times + info
                    "unicycle1_v0/parallelpark_0/rrt_to_v0_guess_v0.yaml",
                    "quadrotor_v0/empty_0_easy/idbastar_v0_db_solution_v0.yaml",  # a single state, no actions
                    "quad2d_v0/empty_0/rrt_to_v0_guess_v0.yaml",
                    "car1_v0/kink_0/idbastar_v0_db_solution_v0.yaml"]


def make_quadrotor_compact(n: int = 3000):
    from dynoplan_b200.motions import load_motion_primitives_new
    from dynoplan_b200.robots import robot_factory
    robot = robot_factory("quad3d_v0", [-10] * 3, [10] * 3)
    lib = load_motion_primitives_new(os.path.join(REF, "dynomotions", SYSTEMS["quadrotor_v0"][0]), robot, n,
                                     add_noise=False)
    # undo the loader's quaternion normalisation?  no: it is applied again on load and is idempotent
    # only to rounding -- so store the RAW file values instead
    import msgpack as mp
    with open(os.path.join(REF, "dynomotions", SYSTEMS["quadrotor_v0"][0]), "rb") as f:
        data = mp.unpackb(f.read(), raw=False)["data"][:n]
    T = np.array([len(t["states"]) for t in data], dtype=np.int16)
    first = np.array([t["states"][0] for t in data], dtype=np.float64)
    last = np.array([t["states"][-1] for t in data], dtype=np.float64)
    actions = np.concatenate([np.asarray(t["actions"], dtype=np.float64) for t in data])
    assert np.array_equal(T, lib.T) and actions.shape == lib.actions.shape
    np.savez_compressed(os.path.join(HERE, "motions", f"quadrotor_v0_{n}_compact.npz"), T=T, first=first, last=last,
                        actions=actions)


def make_expand_fixtures():
    """Oracle expansion records on a fixed node batch per system (golden vectors for the
    oracle itself on CPU and for the CUDA path on the GPU)."""
    from dynoplan_b200.motions import load_motion_primitives_new
    from dynoplan_b200.problem import Problem
    from dynoplan_b200.robots import robot_factory
    from dynoplan_b200 import testing
    from oracle import oracle as orc

    cases = {"unicycle1_v0": "bugtrap_0", "unicycle2_v0": "bugtrap_0", "car1_v0": "kink_0",
             "quad2d_v0": "quad_bugtrap", "acrobot_v0": "swing_up_obs", "quadrotor_v0": "window"}
    for system, env in cases.items():
        p = Problem.from_yaml(os.path.join(HERE, "envs", system, env + ".yaml"))
        robot = robot_factory(p.robotType, p.p_lb, p.p_ub)
        lib = load_motion_primitives_new(os.path.join(HERE, "motions", system + ".msgpack"), robot, rand_seed=1)
        delta = testing.DEFAULT_DELTA[p.robotType]
        radius = float(np.float32(.5) * np.float32(delta))
        lo = p.p_lb if robot.translation_invariance else [0, 0]
        hi = p.p_ub if robot.translation_invariance else [0, 0]
        nodes = testing.frontier_nodes(robot, lib, 48, lo, hi, seed=7)
        nodes[0] = p.start
        nodes[1] = p.goal
        m = orc.OracleModel(robot, p.obstacles)
        ol = orc.OracleLibrary(lib)
        out = orc.expand(m, ol, nodes, radius, p.goal, float(delta), 0)
        np.savez_compressed(os.path.join(HERE, f"expand_{system}.npz"), nodes=nodes, goal=np.array(p.goal),
                            radius=radius, goal_radius=float(delta), env=env,
                            row_ptr=out.row_ptr, prim_idx=out.prim_idx, delta_dist=out.delta_dist,
                            status=out.status, x_final=out.x_final, chosen_index=out.chosen_index,
                            cost_motion=out.cost_motion, jump_lbt=out.jump_lbt, h=out.h, band=out.band)
        st = np.bincount(out.status, minlength=4)
        print(system, env, "edges", len(out.prim_idx), "status counts", st.tolist())


if __name__ == "__main__":
    main()
    make_quadrotor_compact()
