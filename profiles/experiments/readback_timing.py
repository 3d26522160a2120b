"""Where an end-to-end step spends its time: CLExecute, CLReadImageAsync, CLReadImageWait(1), against the
blocking read-backs, on the 640x480 and 1080p workloads.  Run on a GPU box."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
import clpathtracer_b200 as cl
from clpathtracer_b200 import scenes

L = cl.lib()
out = {}
for name, grid, w, h, spp, depth in (("c1", 22, 640, 480, 1, 2), ("1080p_4spp_100k", 224, 1920, 1080, 4, 2)):
    v, c, n = scenes.heightfield(grid, name == "c1")
    scene = cl.build_kd_sah(v, c, n, intersect_cost=1.0, empty_bonus=0.9)
    cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
    r = cl.Renderer(device=0)
    r.set_meshes(scene); r.set_camera_matrix(cam)
    r.set_params(mode=1, depth=depth, spp=spp, seed=0, flags=cl.FLAG_JITTER)
    r.create_image(w, h)
    pinned8 = [torch.empty((h, w, 4), dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    pinned32 = torch.empty((h, w, 4), dtype=torch.float32).pin_memory().numpy()
    pageable8 = np.empty((h, w, 4), dtype=np.uint8)
    for _ in range(5):
        r.execute()
    res = {}
    def timed(label, fn, reps=50):
        ts = []
        for k in range(reps):
            t0 = time.perf_counter(); fn(k); ts.append((time.perf_counter() - t0) * 1e3)
        res[label] = {"p50_ms": round(float(np.median(ts)), 4), "max_ms": round(float(np.max(ts)), 4)}
    timed("execute only", lambda k: r.execute())
    timed("execute + CLReadImage(float4, pinned)", lambda k: (r.execute(), r.read_image(pinned32)))
    timed("execute + CLReadImageRGBA8(pinned)", lambda k: (r.execute(), r.read_image_rgba8(pinned8[0])))
    timed("execute + CLReadImageRGBA8(pageable)", lambda k: (r.execute(), r.read_image_rgba8(pageable8)))
    parts = {"execute": [], "async": [], "wait1": []}
    def step(k):
        t0 = time.perf_counter(); r.execute(); t1 = time.perf_counter()
        r.read_image_async(pinned8[k & 1]); t2 = time.perf_counter()
        r.read_wait(1); t3 = time.perf_counter()
        parts["execute"].append((t1 - t0) * 1e3); parts["async"].append((t2 - t1) * 1e3); parts["wait1"].append((t3 - t2) * 1e3)
    timed("execute + CLReadImageAsync(RGBA8) + Wait(1)", step)
    r.read_wait(0)
    res["pipelined parts p50_ms"] = {k: round(float(np.median(v)), 4) for k, v in parts.items()}
    res["pipelined parts max_ms"] = {k: round(float(np.max(v)), 4) for k, v in parts.items()}
    res["kernel_ms"] = round(r.kernel_ms(), 4)
    out[name] = res
    print(name, json.dumps(res, indent=1), flush=True)
    r.close()
json.dump(out, open("gpurun_out/readback_timing.json", "w"), indent=1)
