"""Bench frame (c3) on ONE GPU: render-kernel ms of EVERY rank's share under 8-way row-tile sharding, for
several tile heights.  The slowest share is what the 8-GPU frame takes.  Run on a GPU box."""
import os, sys, json, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
import clpathtracer_b200 as cl
from clpathtracer_b200 import scenes
L = cl.lib()
v, c, n = scenes.heightfield(707, False)
scene = cl.build_kd_sah(v, c, n, nbins=0, intersect_cost=1.0, empty_bonus=0.9)
w, h = 1920, 1080
cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
r = cl.Renderer(device=0)
r.set_meshes(scene); r.set_camera_matrix(cam)
r.set_params(mode=1, depth=5, spp=64, seed=0, flags=cl.FLAG_JITTER)
out = {}
for tile_rows in (4, 8, 16):
    per_rank = []
    for rank in range(8):
        L.CLSetTileShard(rank, 8, tile_rows)
        r.create_image(w, h)
        ms = []
        for k in range(5):
            L.CLFlushL2()
            r.execute()
            if k >= 2: ms.append(L.CLLastKernelMs())
        per_rank.append(round(float(np.mean(ms)), 4))
    out[f"tile_rows {tile_rows}"] = {"per_rank_ms": per_rank, "max": max(per_rank), "mean": round(float(np.mean(per_rank)), 4)}
    print(tile_rows, out[f"tile_rows {tile_rows}"], flush=True)
L.CLSetTileShard(0, 1, 8)
json.dump(out, open("gpurun_out/shard_balance.json", "w"), indent=1)
