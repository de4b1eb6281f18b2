"""Multi-GPU plumbing: the path shards by samples-per-pixel.

Rank g of G renders the full frame for samples [g*N/G, (g+1)*N/G) -- disjoint PCG32 streams, because the
stream id contains the sample index -- and the float4 accumulators (sum of finite samples, count) are
combined by ONE sum all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).  The mean and the
luminance clamp (main.cpp:168-173) come after the reduction, so the result does not depend on G beyond
float summation order.
"""
import torch
import torch.distributed as dist


def shard_range(n_samples, rank, world):
    """Half-open sample range of `rank`; ranges tile [0, n) exactly, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_samples * rank // world, n_samples * (rank + 1) // world


def allreduce_accumulator(acc):
    """In-place sum of the accumulator across all ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def finalize_torch(acc, max_luminance=1000.0):
    """Reference finalisation on a torch tensor (used by the CPU tests; the GPU path uses mrt_gpu_finalize_device)."""
    cnt = acc[..., 3:4]
    color = torch.where(cnt > 0, acc[..., :3] / cnt, torch.zeros_like(acc[..., :3]))
    c = torch.tensor([0.212655, 0.715158, 0.072187], dtype=acc.dtype, device=acc.device)
    prod = color * c
    lum = (prod[..., 0] + prod[..., 1]) + prod[..., 2]
    over = lum > max_luminance
    scale = torch.where(over, max_luminance / lum, torch.ones_like(lum))
    return torch.where(over[..., None], color * scale[..., None], color)
