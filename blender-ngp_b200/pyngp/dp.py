"""Host-side arithmetic of ray-sharded data parallelism: bindings of the library's own functions (csrc/testbed.cu uses the same ones in
train()), callable without a GPU -- tests/test_data_parallel.py runs them under torch.distributed's gloo backend with two processes."""
import ctypes as C


def next_multiple(v, d):
    return (v + d - 1) // d * d


def shard(rank, world, rays_per_batch):
    """(ray_offset, n_rays_global) of a rank: ngpb_ray_shard."""
    import pyngp
    off, tot = C.c_uint32(), C.c_uint32()
    pyngp.lib().ngpb_ray_shard(int(rank), int(world), int(rays_per_batch), C.byref(off), C.byref(tot))
    return off.value, tot.value


def next_rays_per_batch(rays_per_batch, batch, global_compacted, world):
    """ngpb_next_rays_per_batch: NerfCounters::update_after_training (src/testbed_nerf.cu:2890-2891) on the per-rank average of the all-reduced compacted count."""
    import pyngp
    return int(pyngp.lib().ngpb_next_rays_per_batch(int(rays_per_batch), int(batch), int(global_compacted), int(world)))
