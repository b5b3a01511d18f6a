"""Host-side multi-GPU logic on CPU: sample-range partitioning, and the framebuffer sum over a 2-rank gloo group."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cornelis_b200.sharding import sample_range, sum_framebuffers  # noqa: E402


def test_sample_ranges_partition_exactly():
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 7, 64, 4096, 16384):
            covered = []
            for rank in range(world):
                first, count, total = sample_range(rank, world, spp, "strong")
                assert total == spp
                covered += list(range(first, first + count))
            assert covered == list(range(spp))
            firsts = [sample_range(r, world, spp, "weak") for r in range(world)]
            assert [f[0] for f in firsts] == [r * spp for r in range(world)]
            assert all(f[1] == spp and f[2] == world * spp for f in firsts)
    with pytest.raises(ValueError):
        sample_range(2, 2, 16)
    with pytest.raises(ValueError):
        sample_range(0, 2, 16, "diagonal")


def _fake_accumulators(first, count, npix=257):
    """A stand-in for a rank's accumulation image: the sum over its samples of a per-(pixel, sample) value."""
    s = np.arange(first, first + count, dtype=np.float64)[None, :]
    p = np.arange(npix, dtype=np.float64)[:, None]
    contrib = np.sin(0.37 * p + 1.3 * s) ** 2
    return torch.from_numpy(np.stack([contrib.sum(1), (2 * contrib).sum(1), (3 * contrib).sum(1), np.full(npix, count)], 1)
                            .astype(np.float32)).reshape(-1)


def _worker(rank, world, port, scaling, spp, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count, total = sample_range(rank, world, spp, scaling)
    fb = _fake_accumulators(first, count)
    sum_framebuffers(fb)
    if rank == 0:
        torch.save((fb, total), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scaling", ["weak", "strong"])
def test_two_rank_gloo_sum_equals_single_rank(tmp_path, scaling):
    spp, world = 48, 2
    out = tmp_path / "fb.pt"
    port = 29600 + (os.getpid() % 200) + (0 if scaling == "weak" else 1)
    mp.spawn(_worker, args=(world, port, scaling, spp, str(out)), nprocs=world, join=True)
    fb, total = torch.load(out)
    expect = _fake_accumulators(0, total)
    assert total == (spp * world if scaling == "weak" else spp)
    assert torch.allclose(fb, expect, rtol=1e-5, atol=1e-4)
    assert torch.equal(fb.reshape(-1, 4)[:, 3], torch.full((257,), float(total)))
