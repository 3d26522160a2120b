"""How full the resident warps are during one rank's share of the bench frame (c3) on ONE GPU, and how long the
longest claim is: CLPT_VERBOSE=3 makes CLExecute print it (clstate.cu).  Run on a GPU box."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
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
for nranks, rank in ((1, 0), (8, 0), (8, 3), (8, 7)):
    for g in ("1", "2", "4"):
        os.environ["CLPT_WARPS_PER_PIXEL"] = g
        L.CLSetTileShard(rank, nranks, 8)
        r.create_image(w, h)
        os.environ["CLPT_VERBOSE"] = "0"
        for k in range(4):
            r.execute()
        sys.stderr.write(f"--- shard {rank}/{nranks} warps per pixel {g}\n"); sys.stderr.flush()
        os.environ["CLPT_VERBOSE"] = "3"
        r.execute()
        os.environ["CLPT_VERBOSE"] = "0"
r.close()
