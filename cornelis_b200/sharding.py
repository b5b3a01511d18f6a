"""Sample-index sharding across ranks (one process per GPU).

Pixel-samples are independent (the estimator is a plain mean, reference src/Render.cpp:245-250) and the device
random numbers are keyed by the GLOBAL sample index, so rank r of R renders a contiguous range of sample indices of
every pixel into its own accumulation image, and the images are summed once at the end — the only exchange step.
"""
from __future__ import annotations


def sample_range(rank: int, world: int, spp: int, scaling: str = "weak"):
    """(first_sample, sample_count, total_spp) for this rank.

    weak:   every rank renders `spp` samples; the image ends up with world * spp samples per pixel.
    strong: `spp` samples are split as evenly as possible; ranges are contiguous and disjoint.
    """
    if not (0 <= rank < world) or spp <= 0:
        raise ValueError("bad rank/world/spp")
    if scaling == "weak":
        return rank * spp, spp, world * spp
    if scaling == "strong":
        first = spp * rank // world
        last = spp * (rank + 1) // world
        return first, last - first, spp
    raise ValueError(f"unknown scaling {scaling!r}")


def sum_framebuffers(tensor, group=None):
    """In-place sum of the per-rank accumulation images (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor
