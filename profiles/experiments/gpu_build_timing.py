"""Device-side kd build (CLBuildMeshes) of the heightfields: device time of build and re-layout,
tree statistics, and the frame time of the bench camera on the device-built tree next to the
host SAH trees.  Under `ncu --metrics gpu__time_duration.sum` the launch list gives the time
per kernel of one build.  Run on a GPU box."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import clpathtracer_b200 as cl
from clpathtracer_b200 import scenes

grids = [int(x) for x in sys.argv[1:]] or [158, 707]
r = cl.Renderer(device=0)
out = {}
for g in grids:
    v, c, n = scenes.heightfield(g, False)
    for _ in range(2):
        t0 = time.perf_counter()
        r.build_meshes(v, c, n)
        wall = (time.perf_counter() - t0) * 1e3
    b, p = r.build_ms()
    nodes, refs, levels = (cl.C.c_int(), cl.C.c_int(), cl.C.c_int())
    cl.lib().CLBuildStats(cl.C.byref(nodes), cl.C.byref(refs), cl.C.byref(levels))
    res = {"triangles": len(c) // 3, "device_build_ms": round(b, 3), "device_relayout_ms": round(p, 3),
           "CLBuildMeshes_wall_ms": round(wall, 3), "nodes": nodes.value, "tri_refs": refs.value, "levels": levels.value}
    if "--frames" in os.environ.get("CLPT_BUILD_TIMING_OPTS", ""):
        w, h = 1920, 1080
        cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
        r.set_camera_matrix(cam)
        r.set_params(mode=1, depth=2, spp=4, seed=0, flags=cl.FLAG_JITTER)
        r.create_image(w, h)
        for _ in range(3):
            r.execute()
        res["frame_ms_device_tree"] = round(min(_ms(r) if False else (r.execute(), r.kernel_ms())[1] for _ in range(5)), 4)
        for label, kw in (("host_sah_binned", dict()), ("host_sah_exact", dict(nbins=0))):
            t0 = time.perf_counter()
            s = cl.build_kd_sah(v, c, n, intersect_cost=1.0, empty_bonus=0.9, **kw)
            res[label + "_build_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
            res[label + "_tri_refs"] = s.stats()["leaf_tri_refs"]
            r.set_meshes(s)
            for _ in range(3):
                r.execute()
            res["frame_ms_" + label] = round(min((r.execute(), r.kernel_ms())[1] for _ in range(5)), 4)
    out[f"hf{g}"] = res
    print(g, res, flush=True)
r.close()
json.dump(out, open("gpurun_out/gpu_build_timing.json", "w"), indent=1)
