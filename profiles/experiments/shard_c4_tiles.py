"""Config 4 (4K, 1 spp, depth 5, 10M triangles, progressive) on ONE GPU: the render-kernel time of one
rank's share of the frame under row-tile sharding (CLSetTileShard without a communicator), for several
tile heights.  Each share should cost 1/N of the frame; what it costs more is what caps multi-GPU scaling
of this config.  Run on a GPU box:  python profiles/experiments/shard_c4_tiles.py [grid]"""
import os, sys, json, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch
import clpathtracer_b200 as cl
from clpathtracer_b200 import scenes
L = cl.lib()
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
v, c, n = scenes.heightfield(grid, False)
scene = cl.build_kd_sah(v, c, n, nbins=0, intersect_cost=1.0, empty_bonus=0.9)
w, h = 3840, 2160
cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
r = cl.Renderer(device=0)
r.set_meshes(scene); r.set_camera_matrix(cam)
r.set_params(mode=1, depth=5, spp=1, seed=0, flags=cl.FLAG_JITTER | cl.FLAG_ACCUMULATE)
out = {}
for nranks, rank, tile_rows in ((1, 0, 8), (2, 0, 8), (2, 0, 32), (2, 0, 128), (8, 0, 8), (8, 3, 8), (8, 0, 32), (8, 3, 32),
                                (8, 0, 128), (8, 3, 128), (8, 3, 268)):
    r.create_image(w, h)
    L.CLSetTileShard(rank, nranks, tile_rows)
    ms = []
    for k in range(6):
        if "--flush" in sys.argv: L.CLFlushL2()
        r.execute()
        if k >= 2: ms.append(L.CLLastKernelMs())
    key = f"shard {rank}/{nranks} tile_rows {tile_rows}"
    out[key] = round(float(np.mean(ms)), 4)
    print(f"{key}: {np.mean(ms):.4f} ms  (x{nranks} = {np.mean(ms) * nranks:.2f})", flush=True)
L.CLSetTileShard(0, 1, 8)
json.dump(out, open("gpurun_out/shard_c4_tiles.json", "w"), indent=1)
