"""Multi-GPU plumbing for the P-frame path: GOPs shard over ranks with no data-path collective; the only exchange
is the reduction of the per-rank bit / distortion sums.  One process per GPU
(torch.distributed, NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_units(n_units: int, rank: int, world: int) -> List[int]:
    """Static round-robin of work units (GOPs / sequences) over ranks: unit u goes to rank u % world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_units, world))


def reduce_stats(stats: torch.Tensor) -> torch.Tensor:
    """In-place SUM over ranks of the [bits_y, bits_z, sq_err] x frames accumulator the entropy / synthesis-tail
    kernels wrote (fp64). 3*T doubles: latency-bound, no packing kernel, same stream as the producers."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def summarize(stats_sum: torch.Tensor, n_frames_total: int, pixels_per_frame: int) -> Tuple[float, float]:
    """(mean bpp, PSNR of the mean MSE) over all frames of all ranks, as evalSTEM.py:131-147 aggregates them."""
    s = stats_sum.double().cpu()
    bits = float(s[0].sum() + s[1].sum())
    sq = float(s[2].sum())
    bpp = bits / (n_frames_total * pixels_per_frame)
    mse = sq / (n_frames_total * pixels_per_frame * 3)
    psnr = float("inf") if mse == 0 else -10.0 * torch.log10(torch.tensor(mse)).item()
    return bpp, psnr
