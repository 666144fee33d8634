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


# ---- progressive accumulation, checkpoint / resume, dynamic chunk dispatch ----
#
# The reference's ThreadWorkQueue (code/platform.h:307-339, code/macos_main.mm:165-240) hands 32x32
# pixel tiles to 9 threads and has no way to stop and continue a render.  Here the unit of work is
# a chunk index -- chunk c of every pixel is an independent set of sample streams seeded by
# ort_stream_seed(base_seed, pixel, c) -- and the framebuffer is a sum of 64-bit integers, so chunks
# may be rendered in any order, by any rank, in any number of sessions: the image is bit-identical.

def _fingerprint(params):
    keys = ("output_width", "output_height", "tile_min_x", "tile_min_y", "tile_one_past_max_x", "tile_one_past_max_y",
            "ray_per_pixel_count", "chunk_spp", "base_seed", "russian_roulette_value")
    return [float(getattr(params, k)) for k in keys]


class ProgressiveRender:
    """Accumulates a frame chunk by chunk; can be saved and resumed.

    render_accum(params, accum)  adds chunks [params.chunk_begin, params.chunk_end) into `accum`
                                 (int64 [H, W, 4] tensor); on a GPU: Scene.render_accumulate_device
    """

    def __init__(self, params, accum, render_accum):
        import numpy as np
        self.params, self.accum, self.render_accum = params, accum, render_accum
        self.n_chunks = chunk_count(params.ray_per_pixel_count, params.chunk_spp)
        self.done = np.zeros(self.n_chunks, bool)
        accum.zero_()

    def pending(self):
        return [int(c) for c in range(self.n_chunks) if not self.done[c]]

    def samples_done(self):
        """samples per pixel accumulated so far (the divisor of a preview image)"""
        spp, cs = self.params.ray_per_pixel_count, min(self.params.chunk_spp or self.params.ray_per_pixel_count,
                                                       self.params.ray_per_pixel_count)
        return int(sum(min(cs, spp - c * cs) for c in range(self.n_chunks) if self.done[c]))

    def render_chunks(self, begin, end):
        """renders chunk indices [begin, end); chunks already done are skipped"""
        c = begin
        while c < end:
            if self.done[c]:
                c += 1
                continue
            e = c
            while e < end and not self.done[e]:
                e += 1
            self.params.chunk_begin, self.params.chunk_end = c, e
            self.render_accum(self.params, self.accum)
            self.done[c:e] = True
            c = e

    def step(self, n=1):
        """renders the next `n` pending chunks; returns how many remain"""
        todo = self.pending()[:n]
        for c in todo:
            self.render_chunks(c, c + 1)
        return int((~self.done).sum())

    def save(self, path):
        import numpy as np
        np.savez(path, accum=self.accum.cpu().numpy(), done=self.done, fingerprint=np.array(_fingerprint(self.params)))

    def load(self, path):
        """continues from a checkpoint written by save(); refuses one made with other parameters"""
        import numpy as np
        import torch
        with np.load(path) as z:
            if list(z["fingerprint"]) != _fingerprint(self.params) or z["done"].shape != self.done.shape:
                raise ValueError("checkpoint was made with different render parameters")
            self.accum.copy_(torch.from_numpy(z["accum"]).to(self.accum.device))
            self.done = z["done"].copy()
        return self


def pull_chunks(store, n_chunks, batch=1, key="ort_next_chunk"):
    """Dynamic dispatch across ranks: yields disjoint [begin, end) chunk ranges taken from a counter
    in the process group's key-value store (`store.add` is atomic), until [0, n_chunks) is used up.
    Ranks that render faster simply come back for more -- the multi-process form of the reference's
    tile queue."""
    while True:
        end = int(store.add(key, batch))
        begin = end - batch
        if begin >= n_chunks:
            return
        yield begin, min(end, n_chunks)


def render_dynamic(render_accum, resolve, params, accum, store, batch=1, group=None, dst=0, key="ort_next_chunk"):
    """render_sharded with chunk ranges pulled from a shared counter instead of a fixed split"""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_chunks = chunk_count(params.ray_per_pixel_count, params.chunk_spp)
    accum.zero_()
    mine = []
    for b, e in pull_chunks(store, n_chunks, batch, key):
        params.chunk_begin, params.chunk_end = b, e
        render_accum(params, accum)
        mine.append((b, e))
    if world > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return (resolve(accum) if rank == dst else None), mine
