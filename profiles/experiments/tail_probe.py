"""Where the tail of an as-shipped frame sits (1080p, 1 ray per pixel, 1M triangles, the reference's own
DEPTH-15 tree, engine 2): kernel ms by lanes per ray (frame-wide), and per block row the summed and the longest
warp-tile durations (CLPT_VERBOSE=4 prints them to stderr).  Run on a GPU box."""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import clpathtracer_b200 as cl
from clpathtracer_b200 import scenes

r = cl.Renderer(device=0)
w, h = 1920, 1080
cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
v, c, nn = scenes.heightfield(int(sys.argv[1]) if len(sys.argv) > 1 else 707, False)
r.set_meshes(cl.build_kd(v, c, nn))
r.set_camera_matrix(cam)
r.set_params(mode=cl.MODE_NORMAL, depth=2)
r.create_image(w, h)
cl.lib().CLSetEngine(2)
for k in (1, 2, 4, 8, 16, 32):
    os.environ["CLPT_LANES_PER_RAY"] = str(k)
    for _ in range(3):
        r.execute()
    ms = []
    for _ in range(5):
        r.execute()
        ms.append(r.kernel_ms())
    print("lanes per ray", k, "kernel ms", round(min(ms), 4), flush=True)
    if k in (2, 8, 32):
        os.environ["CLPT_VERBOSE"] = "4"
        sys.stderr.write(f"--- lanes per ray {k}\n")
        sys.stderr.flush()
        r.execute()
        os.environ["CLPT_VERBOSE"] = "0"
r.close()
