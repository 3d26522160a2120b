"""Render-kernel time of one rank's share of the bench frame on ONE GPU (CLSetTileShard without a
communicator), with the claim-direction heuristic off/on ($CLPT_ROW_ORDER).  Source of
profiles/r01_experiments.json: frame_tail_and_claim_direction.  Run on a GPU box:
    python profiles/experiments/shard_kernel_times.py
"""
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
for nranks in (1, 8):
    for rank in sorted({0, nranks - 1}):
        L.CLSetTileShard(rank, nranks, 8)
        r.create_image(w, h)
        for rev in (0, 1):  # 0 = screen order, 1 = claim direction
            os.environ["CLPT_ROW_ORDER"] = str(rev)
            ms = []
            for k in range(6):
                L.CLFlushL2()
                r.execute()
                if k >= 3: ms.append(L.CLLastKernelMs())
            out[f"shard {rank}/{nranks} ordered {rev}"] = round(float(np.mean(ms)), 4)
            print(f"shard {rank}/{nranks} ordered {rev}: {np.mean(ms):.4f} ms  (x{nranks} = {np.mean(ms)*nranks:.2f})", flush=True)
json.dump(out, open("gpurun_out/row_order_exp.json", "w"), indent=1)
