"""The boundary where the reference calls it, and the multi-GPU side of the C ABI.

  * ort_tiled_raytrace_bvh from several host threads at once on disjoint tiles -- how the reference's
    thread callback would use it (code/macos_main.mm:150-163, 574-598: nine workers) -- gives, bit for bit,
    the image of one whole-image call, and returns real primitive-test tallies (ray.cpp:1173).
  * OrtMulti / OrtProgress (ort_multi_*, ort_progress_*): the frame split over the devices of this process,
    summed by one kernel over NVLink peer memory; progressive accumulation, dynamic dispatch, checkpoints.
    On a 1-GPU box the same entry points run with one device; with >= 2 visible GPUs the multi-device tests
    run for real (N-GPU image == 1-GPU image bit for bit).
  * the per-process form: CUDA IPC handles of the framebuffers opened by the root (what bench.py does under
    torchrun), tested with two processes that share GPU 0.
"""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

ROOT = ol.ROOT
SCN = os.path.join(ol.SCENES_DIR, "box_spheres.scn")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_reference_entry_point_from_many_threads(ort, testscene_host):
    W, H, SPP = 480, 270, 4
    hs = testscene_host
    sc = ort.Scene(hs.world, hs.root, 0)
    whole = np.zeros((H, W, 3), np.float32)
    series = np.array([987654321], np.uint32)
    work_whole = sc.tiled_raytrace_bvh(hs.camera, whole, W, H, 0, 0, W, H, series, SPP, 0.8)
    # 6 x 5 = 30 ragged tiles handed to 8 threads, all writing into ONE shared output buffer
    xs = [0, 81, 160, 243, 320, 401, W]; ys = [0, 55, 108, 163, 217, H]
    tiles = [(xs[i], ys[j], xs[i + 1], ys[j + 1]) for j in range(5) for i in range(6)]
    canvas = np.full((H, W, 3), -1.0, np.float32)
    works, errors = [0] * len(tiles), []
    nxt = [0]
    lock = threading.Lock()

    def worker():
        try:
            while True:
                with lock:
                    k = nxt[0]; nxt[0] += 1
                if k >= len(tiles):
                    return
                x0, y0, x1, y1 = tiles[k]
                s = np.array([987654321], np.uint32)          # every tile: its own RandomSeries, same state
                works[k] = sc.tiled_raytrace_bvh(hs.camera, canvas, W, H, x0, y0, x1, y1, s, SPP, 0.8)
        except Exception as e:            # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker) for _ in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert np.array_equal(bits(canvas), bits(whole))
    # the work counter is the number of primitive tests, as the reference's return value: additive over tiles
    assert work_whole > W * H * SPP and sum(works) == work_whole
    sc.close()


def test_render_and_raycast_interleaved_from_threads(ort, testscene_host):
    """different entry points of one handle from different threads: serialised by the handle's lock"""
    hs = testscene_host
    sc = ort.Scene(hs.world, hs.root, 0)
    P = ort.default_params(160, 90, 6, chunk_spp=2)
    hs2 = ort.HostScene.load(os.path.join(ol.DATA_DIR, "testscene.scn"), ol.DATA_DIR, 160, 90)
    want_img, _ = sc.render(hs2.camera, P)
    o, d = ol.make_primary_rays(hs2.camera_array(), 320, 180)
    want_hits = sc.raycast_batch(o, d)
    out, errors = {}, []

    def render(i):
        try:
            out["img%d" % i], _ = sc.render(hs2.camera, ort.default_params(160, 90, 6, chunk_spp=2))
        except Exception as e:            # noqa: BLE001
            errors.append(e)

    def cast(i):
        try:
            out["hit%d" % i] = sc.raycast_batch(o, d)
        except Exception as e:            # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=f, args=(i,)) for i in range(3) for f in (render, cast)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(3):
        assert np.array_equal(bits(out["img%d" % i]), bits(want_img))
        assert np.array_equal(out["hit%d" % i]["rank"], want_hits["rank"]) and np.array_equal(bits(out["hit%d" % i]["t"]), bits(want_hits["t"]))
    sc.close()


# ------------------------------------------------------------------------------------------- OrtMulti
def _devices(ort, want):
    import torch
    return list(range(min(want, torch.cuda.device_count(), ort.device_count())))


@pytest.mark.parametrize("want", [1, 2, 4])
def test_multi_render_equals_single_gpu_image(ort, want):
    devs = _devices(ort, want)
    if len(devs) < want:
        pytest.skip("needs %d GPUs, %d visible" % (want, len(devs)))
    W, H, SPP, CH = 200, 120, 22, 3                        # 8 chunks, the last one ragged
    hs = ort.HostScene.load(SCN, ol.SCENES_DIR, W, H)
    scenes = [ort.Scene(hs.world, hs.root, d) for d in devs]
    single, st1 = scenes[0].render(hs.camera, ort.default_params(W, H, SPP, chunk_spp=CH))
    m = ort.Multi(scenes)
    n, direct = m.device_count()
    assert n == want
    img, st = m.render(hs.camera, ort.default_params(W, H, SPP, chunk_spp=CH))
    assert np.array_equal(bits(img), bits(single))          # bit-identical for any number of GPUs
    assert st["samples"] == W * H * SPP == st1["samples"] and st["rays"] == st1["rays"]
    # whole images only, bad arguments reported
    Q = ort.default_params(W, H, SPP, chunk_spp=CH); Q.tile_min_x = 8
    with pytest.raises(ort.OrtError):
        m.render(hs.camera, Q)
    m.close()
    for sc in scenes:
        sc.close()


@pytest.mark.parametrize("want", [1, 2])
def test_progress_dynamic_dispatch_checkpoint_resume(ort, tmp_path, want):
    devs = _devices(ort, want)
    if len(devs) < want:
        pytest.skip("needs %d GPUs, %d visible" % (want, len(devs)))
    W, H, SPP, CH = 160, 90, 14, 4                          # 4 chunks: 4 + 4 + 4 + 2 samples
    hs = ort.HostScene.load(SCN, ol.SCENES_DIR, W, H)
    scenes = [ort.Scene(hs.world, hs.root, d) for d in devs]
    whole, _ = scenes[0].render(hs.camera, ort.default_params(W, H, SPP, chunk_spp=CH))
    m = ort.Multi(scenes)
    pr = ort.Progress(m, hs.camera, ort.default_params(W, H, SPP, chunk_spp=CH))
    assert pr.state() == dict(chunks_done=0, chunks_total=4, spp_done=0)
    with pytest.raises(ort.OrtError):
        pr.resolve(W, H)                                     # nothing rendered yet
    st = pr.render(1)
    assert pr.state() == dict(chunks_done=1, chunks_total=4, spp_done=4) and st["samples"] == W * H * 4
    # the image so far = the 4-spp image of chunk 0 alone
    part = pr.resolve(W, H)
    Q = ort.default_params(W, H, SPP, chunk_spp=CH); Q.chunk_begin, Q.chunk_end = 0, 1
    import torch
    acc = torch.zeros((H, W, 4), dtype=torch.int64, device="cuda:%d" % devs[0])
    rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:%d" % devs[0])
    with torch.cuda.device(devs[0]):
        scenes[0].render_accumulate_device(hs.camera, Q, acc.data_ptr())
        scenes[0].accum_resolve_device(acc.data_ptr(), W, H, 4, rgb.data_ptr())
        torch.cuda.synchronize()
    assert np.array_equal(bits(part), bits(rgb.cpu().numpy()))
    ck = str(tmp_path / "frame.ortprog")
    pr.save(ck)
    pr.close()
    # a new session continues from the file; every device pulls chunks until none is left
    pr2 = ort.Progress(m, path=ck)
    assert pr2.state() == dict(chunks_done=1, chunks_total=4, spp_done=4)
    pr2.render(0)
    assert pr2.state() == dict(chunks_done=4, chunks_total=4, spp_done=SPP)
    assert pr2.render(0)["samples"] == 0                     # nothing left: idempotent
    got = pr2.resolve(W, H)
    assert np.array_equal(bits(got), bits(whole))
    pr2.close()
    # a file that is not a checkpoint is refused, not crashed on
    bad = tmp_path / "bad.ortprog"
    bad.write_bytes(b"ORTPROG1" + b"\x00" * 40)
    with pytest.raises(ort.OrtError):
        ort.Progress(m, path=str(bad))
    m.close()
    for sc in scenes:
        sc.close()


def test_reduce_resolve_kernel_on_local_peers(ort):
    """the peer-sum kernel with three framebuffers on one device (pointers instead of NVLink peers):
    float3 pixels and RGBE words equal the single-call image / the host encoder"""
    import torch
    W, H, SPP, CH = 120, 70, 12, 2
    hs = ort.HostScene.load(SCN, ol.SCENES_DIR, W, H)
    sc = ort.Scene(hs.world, hs.root, 0)
    whole, _ = sc.render(hs.camera, ort.default_params(W, H, SPP, chunk_spp=CH))
    bufs = [sc.accum_alloc_device(W, H) for _ in range(3)]
    for r, (b, e) in enumerate([(0, 1), (1, 4), (4, 6)]):
        P = ort.default_params(W, H, SPP, chunk_spp=CH); P.chunk_begin, P.chunk_end = b, e
        sc.render_accumulate_device(hs.camera, P, bufs[r])
    rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    rgbe = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
    sc.accum_reduce_resolve_device(bufs[0], bufs[1:], W, H, SPP, rgb.data_ptr(), rgbe.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(bits(rgb.cpu().numpy()), bits(whole))
    want = np.array([[ort.v3_to_rgbe(whole[H - 1 - r, x]) for x in range(W)] for r in range(0, H, 7)], np.uint32)
    assert np.array_equal(rgbe.cpu().numpy().view(np.uint32)[::7], want)
    # no peers: a plain resolve of the (now summed) root buffer
    rgb2 = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    sc.accum_reduce_resolve_device(bufs[0], [], W, H, SPP, rgb2.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(bits(rgb2.cpu().numpy()), bits(whole))
    for b in bufs:
        sc.accum_free_device(b)
    sc.close()


_IPC_CHILD = r"""
import sys, os
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import offline_raytracer_b200 as ort, oracle_lib as ol
W, H, SPP, CH = 120, 70, 12, 2
hs = ort.HostScene.load(%(scn)r, ol.SCENES_DIR, W, H)
sc = ort.Scene(hs.world, hs.root, 0)
buf = sc.accum_alloc_device(W, H)
P = ort.default_params(W, H, SPP, chunk_spp=CH); P.chunk_begin, P.chunk_end = 3, 6
sc.render_accumulate_device(hs.camera, P, buf)
print(sc.accum_ipc_export(buf).hex(), flush=True)
sys.stdin.readline()          # keep the allocation alive until the root has read it
"""


def test_ipc_framebuffer_of_another_process(ort):
    """one process per GPU, as under torchrun: the root opens the other rank's framebuffer through a CUDA IPC
    handle and sums it in place (here both processes share GPU 0)"""
    import torch
    W, H, SPP, CH = 120, 70, 12, 2
    hs = ort.HostScene.load(SCN, ol.SCENES_DIR, W, H)
    sc = ort.Scene(hs.world, hs.root, 0)
    whole, _ = sc.render(hs.camera, ort.default_params(W, H, SPP, chunk_spp=CH))
    child = subprocess.Popen([sys.executable, "-c", _IPC_CHILD % dict(root=ROOT, scn=SCN)],
                             stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
    try:
        line = child.stdout.readline().strip()
        assert len(line) == 128, "child failed: %r" % line
        peer = sc.accum_ipc_open(bytes.fromhex(line))
        mine = sc.accum_alloc_device(W, H)
        P = ort.default_params(W, H, SPP, chunk_spp=CH); P.chunk_begin, P.chunk_end = 0, 3
        sc.render_accumulate_device(hs.camera, P, mine)
        rgb = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
        sc.accum_reduce_resolve_device(mine, [peer], W, H, SPP, rgb.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(bits(rgb.cpu().numpy()), bits(whole))
        sc.accum_ipc_close(peer)
        sc.accum_free_device(mine)
    finally:
        try:
            child.stdin.write("done\n"); child.stdin.flush()
        except Exception:            # noqa: BLE001
            pass
        child.wait(timeout=60)
    sc.close()
