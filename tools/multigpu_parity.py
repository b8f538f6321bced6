#!/usr/bin/env python
"""Parity of the REAL multi-process multi-GPU path, for every robot model.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_parity.py [out.json]

One process per GPU, the library sharded by primitive index, records travelling through the
CUDA-IPC mailboxes (stores over NVLink inside k_edge_tail, one all_reduce as the fence, merge by
ranking) -- exactly what bench.py times at N > 1.  Every rank compares its merged node slice
(all record fields, bit for bit, both mailbox parities) with one unsharded expander over the whole
library on its own GPU and, on a few nodes, with the CPU oracle (bench.parity_sharded).  Exit
code 3 on any mismatch.  bench.py runs the same check on its own workload before it prints."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # robot, primitives (whole library), frontier nodes, obstacles
    ("quad3d_v0", 40_000, 1024, None), ("quad3d_v0", 24_001, 509, 64), ("unicycle1_v0", 12_000, 768, 12),
    ("unicycle2_v0", 12_000, 512, 12), ("car1_v0", 12_000, 512, 12), ("quad2d_v0", 16_000, 512, 12),
    ("acrobot_v0", 8_000, 256, 6),
]


def main():
    import torch
    import torch.distributed as dist
    import bench
    from dynoplan_b200 import sharded, testing
    from dynoplan_b200.expander import GpuExpander
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    results, failed = [], 0
    for k, (robot_type, n_prims, n_nodes, n_obs) in enumerate(CASES):
        lean = k % 2 == 0  # both mailbox formats: with and without final states
        case = testing.synthetic_case(robot_type, n_prims, n_nodes, seed=11, n_obs=n_obs, device=local)
        sb = sharded.shard_bounds(case.library.N, world)
        shard = case.library.slice(sb[rank], sb[rank + 1])
        mine = testing.Case(case.robot, case.obstacles, shard, case.nodes, case.goal, case.radius, case.goal_radius)
        with GpuExpander(case.robot, case.obstacles, shard, device=local, capacity=1 << 23) as ex:
            ex.set_stream(stream.cuda_stream)
            d_nodes, d_goal = torch.from_numpy(case.nodes).cuda(), torch.from_numpy(case.goal).cuda()
            merger = sharded.ShardedExpander(ex, rank, world, prim_offset=sb[rank])
            n0 = ex.expand_count(case.nodes, case.radius, case.goal, case.goal_radius, 0)
            sizes = torch.tensor([n0], device="cuda", dtype=torch.int64)
            dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
            merger.setup_mailbox(n_nodes, int(2 * int(sizes[0]) / world) + 4096, int(2 * int(sizes[0]) / world) + 4096,
                                 lean=lean)
            merger.gather_final_tables(shard.T)
            par = bench.gather_parity(bench.parity_sharded(mine, merger, case.library, d_nodes, d_goal, local), world)
            torch.cuda.synchronize()
            dist.barrier()
        par.update(lean_mailbox=lean, robot=robot_type, n_primitives=n_prims, n_nodes=n_nodes, n_obstacles=len(case.obstacles), gpus=world)
        results.append(par)
        failed += par["mismatch"]
        if rank == 0:
            print(f"{robot_type:13s} G={world} lean={int(lean)} edges={par['edges_compared']:8d} accepted={par['accepted_edges']:8d} "
                  f"mismatch={par['mismatch']}  oracle: {par['per_rank'][0]['oracle']}", flush=True)
    if rank == 0 and len(sys.argv) > 1:
        os.makedirs(os.path.dirname(os.path.abspath(sys.argv[1])), exist_ok=True)
        with open(sys.argv[1], "w") as f:
            json.dump(results, f, indent=1)
    dist.destroy_process_group()
    sys.exit(3 if failed else 0)


if __name__ == "__main__":
    main()
