"""Multi-GPU sharding of the radiance loop: one process per GPU, `torch.distributed` for the plumbing.

The path shards trivially (every (pixel, chunk) sample stream is independent and the scene is
read-only), so each rank holds the whole flattened scene and renders its own range of sample
chunks of the same image into a 64-bit fixed-point framebuffer.  The only exchange step is one
reduce(sum, int64) of that framebuffer to rank 0 (NCCL over NVLink on GPUs; gloo in the CPU
tests).  Integer addition is associative, so the N-GPU image is BIT-IDENTICAL to the 1-GPU image
of the same seed schedule -- there is no summation-order caveat to document.

The reference has no counterpart: its only parallelism is 9 threads pulling 32x32-pixel tiles
from a shared-memory queue (code/macos_main.mm:574-671).
"""


def shard_chunks(n_chunks, world_size, rank):
    """contiguous, balanced split of chunk indices [0, n_chunks) -> [begin, end) of `rank`"""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_chunks, world_size)
    begin = rank * base + min(rank, rem)
    end = begin + base + (1 if rank < rem else 0)
    return begin, end


def chunk_count(spp, chunk_spp):
    chunk_spp = min(chunk_spp or spp, spp)
    return (spp + chunk_spp - 1) // chunk_spp


def render_sharded(render_accum, resolve, params, accum, group=None, dst=0):
    """One frame on `world_size` ranks.

    render_accum(params, accum)  adds this rank's chunk range into `accum` (int64 [H, W, 4] tensor);
                                 on a GPU this is Scene.render_accumulate_device
    resolve(accum)               turns the reduced buffer into pixels on rank `dst`
    Returns resolve()'s value on rank dst, None elsewhere.
    """
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_chunks = chunk_count(params.ray_per_pixel_count, params.chunk_spp)
    params.chunk_begin, params.chunk_end = shard_chunks(n_chunks, world, rank)
    accum.zero_()
    if params.chunk_end > params.chunk_begin:
        render_accum(params, accum)
    if world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    if rank == dst:
        return resolve(accum)
    return None
