#!/usr/bin/env python
"""torchrun form of the multi-GPU identity check: every rank processes its round-robin share of F frames on its own
GPU; rank 0 also processes all F frames alone and compares, bit for bit, the per-frame final clouds (SHA-256), counts
and refined extrinsics gathered from the ranks.  No collective touches the frame path: the gather is the epilogue.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_identity.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from kinectpy_b200 import synth
    from kinectpy_b200.pipeline import FramePipeline, PipelineConfig
    from kinectpy_b200.sharding import frames_for_rank
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mode = synth.NFOV
    F = 2 * world + 1
    depth, tab, T = synth.render_sequence(mode, F, 3)
    Ti = np.stack([synth.perturbed_extrinsic(T[s], 0.3, (3, -3, 3)) if s else T[s] for s in range(3)])
    cfg = PipelineConfig(n_sensors=3, pixels=mode.pixels, n_streams=4)

    def digest(o):
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(o.points).tobytes())
        h.update(np.ascontiguousarray(o.icp_T).tobytes())
        h.update(np.array([o.n_fused, o.n_voxel, o.n_sor, o.n_floor_inliers, o.n_out], np.int64).tobytes())
        return h.hexdigest()

    mine = frames_for_rank(F, rank, world)
    outs = FramePipeline(cfg, tab, T, Ti, device=local).run(np.ascontiguousarray(depth[mine]), want_points=True)
    local_rows = [(int(f), digest(o)) for f, o in zip(mine, outs)]
    rows = [None] * world
    if world > 1:
        dist.all_gather_object(rows, local_rows)
    else:
        rows = [local_rows]
    if rank == 0:
        ref = FramePipeline(cfg, tab, T, Ti, device=local).run(depth, want_points=True)
        want = {f: digest(o) for f, o in enumerate(ref)}
        got = {f: d for r in rows for f, d in r}
        ok = sorted(got) == list(range(F)) and all(got[f] == want[f] for f in range(F))
        print(json.dumps({"check": "multi_gpu_identity", "world": world, "frames": F, "mode": "NFOV C4", "identical": bool(ok),
                          "frames_per_rank": [len(r) for r in rows]}))
        if not ok:
            raise SystemExit(1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
