"""C1-sized jobs: megakernel vs wavefront with pools sized to the job (development)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import offline_raytracer_b200 as ort
data = os.path.join(ROOT, "oracle", "_ref", "data")
for (w, h, spp) in ((480, 270, 16), (960, 540, 16), (1280, 720, 16)):
    hs = ort.HostScene.load(os.path.join(data, "testscene.scn"), data, w, h)
    sc = ort.Scene(hs.world, hs.root, 0)
    def run(kernel):
        P = ort.default_params(w, h, spp, chunk_spp=16, kernel=kernel)
        sc.render(hs.camera, P)
        return min(sc.render(hs.camera, P)[1]["device_ms"] for _ in range(3))
    row = {"job": "%dx%dx%d" % (w, h, spp), "samples_M": w * h * spp / 1e6, "mega_ms": round(run(1), 2)}
    for pools in (1, 2):
        for slots in (128, 256, 512, 1024, 2048, 6144):
            os.environ["ORT_WF_POOLS"] = str(pools); os.environ["ORT_WF_SLOTS"] = str(slots << 10)
            row["wf_p%d_%dKi" % (pools, slots)] = round(run(2), 2)
    print(json.dumps(row), flush=True)
    sc.close()
