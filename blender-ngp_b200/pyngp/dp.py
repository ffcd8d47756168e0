"""Host-side arithmetic of ray-sharded data parallelism, mirrored from csrc/testbed.cu so that it can be exercised
without a GPU (tests/test_data_parallel.py runs it under torch.distributed's gloo backend with two processes)."""


def next_multiple(v, d):
    return (v + d - 1) // d * d


def shard(rank, world, rays_per_batch):
    """(ray_offset, n_rays_global) of a rank: every rank marches `rays_per_batch` rays of a global batch of world * rays_per_batch."""
    return rank * rays_per_batch, world * rays_per_batch


def next_rays_per_batch(rays_per_batch, batch, global_compacted, world):
    """NerfCounters::update_after_training (src/testbed_nerf.cu:2890-2891) on the per-rank average of the all-reduced compacted count,
    in float32 like the reference, so that every rank derives the same value."""
    import numpy as np
    measured = max(1, global_compacted // world)
    r = int(np.float32(np.float32(rays_per_batch) * np.float32(batch)) / np.float32(measured))
    return min(next_multiple(r, 128), 1 << 18)
