"""Multi-GPU layer: frames are independent units (``preprocessing/data.py:35`` is a pure map over
``file_idx``), so ranks take disjoint frame indices and there is NO collective on the per-frame
path.  The only exchange is an optional epilogue that gathers the fixed-size per-frame results
(refined extrinsics, counts) so one rank can write them out -- tens of kilobytes.

One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is plumbing only.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def frames_for_rank(n_frames: int, rank: int, world: int) -> np.ndarray:
    """Round-robin ``f = rank (mod world)``: balanced to within one frame for any F, G."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return np.arange(rank, n_frames, world, dtype=np.int64)


def gather_frame_results(local_frames: Sequence[int], local_T: np.ndarray, local_counts: np.ndarray, n_frames: int,
                         group=None):
    """Epilogue all-gather: returns (T[F,K,4,4], counts[F,C]) on every rank, rows in frame order.

    ``local_T`` float64[n_local,K,4,4]; ``local_counts`` int64[n_local,C].  Ranks may hold different
    numbers of frames (F not a multiple of G): rows are padded to the largest local count.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    K = local_T.shape[1] if local_T.ndim == 4 else 0
    Cn = local_counts.shape[1]
    n_max = (n_frames + world - 1) // world
    dev = torch.device("cuda", torch.cuda.current_device()) if (dist.is_initialized() and dist.get_backend(group) == "nccl") \
        else torch.device("cpu")
    row = 1 + K * 16 + Cn
    buf = torch.full((n_max, row), -1.0, dtype=torch.float64, device=dev)
    for i, f in enumerate(local_frames):
        buf[i, 0] = float(f)
        if K:
            buf[i, 1:1 + K * 16] = torch.from_numpy(np.ascontiguousarray(local_T[i]).reshape(-1)).to(dev)
        buf[i, 1 + K * 16:] = torch.from_numpy(local_counts[i].astype(np.float64)).to(dev)
    if world > 1:
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf, group=group)
        allrows = torch.cat(out, 0).cpu().numpy()
    else:
        allrows = buf.cpu().numpy()
    T = np.zeros((n_frames, K, 4, 4))
    counts = np.zeros((n_frames, Cn), dtype=np.int64)
    seen = np.zeros(n_frames, dtype=bool)
    for r in allrows:
        f = int(r[0])
        if f < 0:
            continue
        seen[f] = True
        if K:
            T[f] = r[1:1 + K * 16].reshape(K, 4, 4)
        counts[f] = np.rint(r[1 + K * 16:]).astype(np.int64)
    if not seen.all():
        raise RuntimeError("gather_frame_results: some frames were not produced by any rank")
    return T, counts
