"""Multi-GPU path (offline_raytracer_b200/dist.py).

CPU part: world_size-2 `gloo` processes shard the sample chunks exactly as the GPU ranks do, with
the CPU oracle standing in for the kernels; the reduced int64 framebuffer must equal the
single-process one bit for bit.  GPU part (-m gpu): the same property through the CUDA path, the
ranks emulated on one device.
"""
import os
import socket
import sys

import numpy as np
import pytest

import oracle_lib as ol

ROOT = ol.ROOT


def test_shard_chunks_partitions_exactly():
    from offline_raytracer_b200.dist import shard_chunks, chunk_count
    for n in (0, 1, 7, 16, 64, 1000):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_chunks(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert chunk_count(256, 16) == 16 and chunk_count(12, 5) == 3 and chunk_count(8, 0) == 1
    with pytest.raises(ValueError):
        shard_chunks(4, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import offline_raytracer_b200 as ort
    from offline_raytracer_b200.dist import render_sharded
    import oracle_lib as ol2
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, SPP, CH = 40, 24, 10, 3            # ragged: chunks of 3,3,3,1 over 2 ranks
    hs = ort.HostScene.load(os.path.join(ol2.SCENES_DIR, "box_spheres.scn"), ol2.SCENES_DIR, W, H)
    osc = ol2.Oracle().scene(hs.world, hs.root)
    P = ol2.default_params(W, H, SPP, chunk_spp=CH)
    accum = torch.zeros((H, W, 4), dtype=torch.int64)

    def render_accum(params, acc):
        osc.render_accum(hs.camera, params, acc.numpy(), threads=2)

    img = render_sharded(render_accum, lambda acc: osc.accum_resolve(acc.numpy(), SPP), P, accum)
    if rank == 0:
        np.save(out_path, img)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process(built, ort, oracle, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "img.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    img2 = np.load(out)
    W, H, SPP, CH = 40, 24, 10, 3
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, W, H)
    osc = oracle.scene(hs.world, hs.root)
    img1, _ = osc.render(hs.camera, ol.default_params(W, H, SPP, chunk_spp=CH))
    assert np.array_equal(img1.view(np.uint32), img2.view(np.uint32))


@pytest.mark.gpu
def test_emulated_ranks_on_one_gpu_equal_single_call(ort):
    """4 'ranks' render their chunk ranges on one device; the summed buffers == one call"""
    import torch
    from offline_raytracer_b200.dist import shard_chunks, chunk_count
    W, H, SPP, CH = 160, 90, 20, 3
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, W, H)
    sc = ort.Scene(hs.world, hs.root, 0)
    whole, _ = sc.render(hs.camera, ort.default_params(W, H, SPP, chunk_spp=CH))
    st = torch.cuda.current_stream().cuda_stream
    total = torch.zeros((H, W, 4), dtype=torch.int64, device="cuda:0")
    n = chunk_count(SPP, CH)
    for rank in range(4):
        acc = torch.zeros((H, W, 4), dtype=torch.int64, device="cuda:0")
        P = ort.default_params(W, H, SPP, chunk_spp=CH)
        P.chunk_begin, P.chunk_end = shard_chunks(n, 4, rank)
        if P.chunk_end > P.chunk_begin:
            sc.render_accumulate_device(hs.camera, P, acc.data_ptr(), stream=st)
        torch.cuda.synchronize()
        total += acc
    rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    sc.accum_resolve_device(total.data_ptr(), W, H, SPP, rgb.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(rgb.cpu().numpy().view(np.uint32), whole.view(np.uint32))
    sc.close()


# ---- progressive accumulation, checkpoint / resume, dynamic dispatch ----

def test_progressive_checkpoint_resume_is_bit_identical(ort, oracle, tmp_path):
    """chunks rendered in two sessions, with a checkpoint in between, sum to the one-shot image"""
    import torch
    from offline_raytracer_b200.dist import ProgressiveRender
    W, H, SPP, CH = 40, 24, 11, 2             # 6 chunks, the last one ragged
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, W, H)
    osc = oracle.scene(hs.world, hs.root)
    want, _ = osc.render(hs.camera, ol.default_params(W, H, SPP, chunk_spp=CH))

    def render_accum(params, acc):
        osc.render_accum(hs.camera, params, acc.numpy(), threads=2)

    pr = ProgressiveRender(ol.default_params(W, H, SPP, chunk_spp=CH), torch.zeros((H, W, 4), dtype=torch.int64), render_accum)
    assert pr.n_chunks == 6 and pr.samples_done() == 0
    pr.render_chunks(4, 6)                   # out of order on purpose
    assert pr.step(1) == 3 and pr.samples_done() == 2 + 1 + 2
    ck = str(tmp_path / "ck.npz")
    pr.save(ck)
    # a new session: fresh accumulator, continue from the file
    pr2 = ProgressiveRender(ol.default_params(W, H, SPP, chunk_spp=CH), torch.zeros((H, W, 4), dtype=torch.int64), render_accum).load(ck)
    assert pr2.pending() == [1, 2, 3]
    while pr2.step(2):
        pass
    assert pr2.samples_done() == SPP
    got = osc.accum_resolve(pr2.accum.numpy(), SPP)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # a checkpoint of other parameters is refused
    other = ProgressiveRender(ol.default_params(W, H, SPP + 1, chunk_spp=CH), torch.zeros((H, W, 4), dtype=torch.int64), render_accum)
    with pytest.raises(ValueError):
        other.load(ck)


def _dyn_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import offline_raytracer_b200 as ort
    from offline_raytracer_b200.dist import render_dynamic
    import oracle_lib as ol2
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    store = dist.distributed_c10d._get_default_store()
    W, H, SPP, CH = 40, 24, 10, 1            # 10 chunks pulled one or two at a time
    hs = ort.HostScene.load(os.path.join(ol2.SCENES_DIR, "box_spheres.scn"), ol2.SCENES_DIR, W, H)
    osc = ol2.Oracle().scene(hs.world, hs.root)
    P = ol2.default_params(W, H, SPP, chunk_spp=CH)
    accum = torch.zeros((H, W, 4), dtype=torch.int64)

    def render_accum(params, acc):
        osc.render_accum(hs.camera, params, acc.numpy(), threads=1)

    img, mine = render_dynamic(render_accum, lambda acc: osc.accum_resolve(acc.numpy(), SPP), P, accum, store, batch=2)
    np.save(os.path.join(out_dir, "ranges%d.npy" % rank), np.array(mine, np.int64).reshape(-1, 2))
    if rank == 0:
        np.save(os.path.join(out_dir, "img.npy"), img)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dynamic_dispatch_covers_every_chunk_once(built, ort, oracle, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_dyn_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ranges = np.concatenate([np.load(str(tmp_path / ("ranges%d.npy" % r))) for r in range(2)])
    covered = sorted(c for b, e in ranges for c in range(b, e))
    assert covered == list(range(10))
    W, H, SPP, CH = 40, 24, 10, 1
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, W, H)
    img1, _ = oracle.scene(hs.world, hs.root).render(hs.camera, ol.default_params(W, H, SPP, chunk_spp=CH))
    assert np.array_equal(img1.view(np.uint32), np.load(str(tmp_path / "img.npy")).view(np.uint32))


@pytest.mark.gpu
def test_progressive_render_on_the_gpu_with_checkpoint(ort, tmp_path):
    import torch
    from offline_raytracer_b200.dist import ProgressiveRender
    W, H, SPP, CH = 160, 90, 14, 4
    hs = ort.HostScene.load(os.path.join(ol.SCENES_DIR, "box_spheres.scn"), ol.SCENES_DIR, W, H)
    sc = ort.Scene(hs.world, hs.root, 0)
    whole, _ = sc.render(hs.camera, ort.default_params(W, H, SPP, chunk_spp=CH))
    st = torch.cuda.current_stream().cuda_stream

    def render_accum(params, acc):
        sc.render_accumulate_device(hs.camera, params, acc.data_ptr(), stream=st)
        torch.cuda.synchronize()

    pr = ProgressiveRender(ort.default_params(W, H, SPP, chunk_spp=CH), torch.zeros((H, W, 4), dtype=torch.int64, device="cuda:0"), render_accum)
    pr.step(1)
    pr.save(str(tmp_path / "ck.npz"))
    pr2 = ProgressiveRender(ort.default_params(W, H, SPP, chunk_spp=CH), torch.zeros((H, W, 4), dtype=torch.int64, device="cuda:0"),
                            render_accum).load(str(tmp_path / "ck.npz"))
    while pr2.step(2):
        pass
    rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    sc.accum_resolve_device(pr2.accum.data_ptr(), W, H, SPP, rgb.data_ptr(), stream=st)
    rgbe = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
    sc.accum_resolve_rgbe_device(pr2.accum.data_ptr(), W, H, SPP, rgbe.data_ptr(), stream=st)
    torch.cuda.synchronize()
    assert np.array_equal(rgb.cpu().numpy().view(np.uint32), whole.view(np.uint32))
    want = np.array([[ort.v3_to_rgbe(whole[H - 1 - r, x]) for x in range(W)] for r in range(0, H, 9)], np.uint32)
    assert np.array_equal(rgbe.cpu().numpy().view(np.uint32)[::9], want)
    sc.close()
