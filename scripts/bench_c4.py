#!/usr/bin/env python
"""BASELINE.json configs[3]: batched loop-closure — 256 synthetic candidate scan pairs (131 072 points each, loop-sized
relative motion), createConstraint(loop = true) semantics: a coarse align (density 0.5 m, correspondence distance 5 m) whose
result is the guess of the fine align (density 0.1 m), PointCloudSensor.cpp:286-292.  The 256 pairs are sharded over the
ranks (strong scaling, no data-path collective); rank 0 gathers the results like the graph-owning thread would.

    python scripts/bench_c4.py                                        1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_c4.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=4, help="distinct scenes per rank, cycled")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--two-calls", action="store_true", help="coarse batch then fine batch instead of s3d_gicp_align_loop_batch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    import slam3d_b200
    from slam3d_b200 import _abi, sharding, synth
    from slam3d_b200._abi import RegistrationParameters

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = slam3d_b200.Context([local_rank])
    lo, hi = sharding.shard_range(args.pairs, rank, world)
    scenes = [synth.scan_pair(seed=5000 + 100 * rank + i, loop=True) for i in range(args.distinct)]
    srcs = [torch.from_numpy(slam3d_b200.as_xyzw(scenes[i % args.distinct][0])).pin_memory() for i in range(hi - lo)]
    tgts = [torch.from_numpy(slam3d_b200.as_xyzw(scenes[i % args.distinct][1])).pin_memory() for i in range(hi - lo)]
    truth = [scenes[i % args.distinct][2] for i in range(hi - lo)]
    coarse = RegistrationParameters.defaults(point_cloud_density=0.5, max_correspondence_distance=5.0, max_translation=5.0)
    fine = RegistrationParameters.defaults(point_cloud_density=0.1, max_translation=5.0)

    def step():
        if args.two_calls:  # the unfused sequence: every scan crosses PCIe twice, and the phases do not overlap
            rc = ctx.gicp_align_batch(srcs, tgts, None, coarse)
            rf = ctx.gicp_align_batch(srcs, tgts, [r.pose() for r in rc], fine)
            return rc, rf
        return ctx.gicp_align_loop_batch(srcs, tgts, None, coarse, fine)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc, rf = step()
    barrier()
    dt = sharding.max_over_ranks(time.perf_counter() - t0, device="cuda") / args.steps
    ok = [r.status == _abi.S3D_OK for r in rf]
    err = []
    for r, T, good in zip(rf, truth, ok):
        if good:
            D = np.linalg.inv(T) @ r.pose()
            err.append(float(np.linalg.norm(D[:3, 3])))
    stats = sharding.gather_to_rank0([(int(sum(ok)), len(ok), float(np.median(err)) if err else None,
                                       float(np.mean([r.outer_iterations for r in rc])), float(np.mean([r.outer_iterations for r in rf])))])
    if rank == 0:
        print(json.dumps({"workload": "BASELINE configs[3]: 256 loop-closure candidate pairs, coarse (0.5 m, 5 m) then fine (0.1 m) GICP, host scans in",
                          "n_gpus": world, "pairs": args.pairs, "api": "two s3d_gicp_align_batch calls" if args.two_calls else "s3d_gicp_align_loop_batch", "ms_per_step": 1e3 * dt, "loop_constraints_per_s": args.pairs / dt,
                          "aligns_per_s": 2 * args.pairs / dt, "accepted": sum(s[0] for s in stats), "of": sum(s[1] for s in stats),
                          "median_translation_error_m_vs_truth": [s[2] for s in stats][:2],
                          "mean_outer_iterations_coarse_fine": [stats[0][3], stats[0][4]]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
