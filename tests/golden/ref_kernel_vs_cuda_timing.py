"""Same B200, same scene, same reference-built DEPTH-15 tree, same mode (as
shipped: 1 ray per pixel, normal colour): the reference's kernel.cl through
NVIDIA's OpenCL against this repository's CUDA kernel.  Run on the GPU box:

    gpurun -- 'python tests/golden/ref_kernel_vs_cuda_timing.py gpurun_out/ref_vs_cuda.json'
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import clpathtracer_b200 as cl  # noqa: E402
from clpathtracer_b200 import scenes  # noqa: E402
from oracle import oracle_py as op  # noqa: E402


def main(out_path):
    ok, what = op.ref_kernel_available()
    if not ok:
        raise SystemExit(what)
    r = cl.Renderer(device=0)
    out = {"device": what, "rows": []}
    w, h = 1920, 1080
    cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
    for n in (22, 224, 707):
        v, c, nn = scenes.heightfield(n, False)
        for label, scene in (("reference tree (DEPTH 15, 25 bins)", cl.build_kd(v, c, nn)),
                             ("SAH tree", cl.build_kd_sah(v, c, nn, intersect_cost=1.0, empty_bonus=0.9))):
            _, ref_ms = op.ref_kernel_render(scene, cam, w, h, repeats=5)
            r.set_meshes(scene)
            r.set_camera_matrix(cam)
            r.set_params(mode=cl.MODE_NORMAL, depth=2)
            r.create_image(w, h)
            per_engine = {}
            for engine in (1, 2, 0):  # lane-per-ray, warp-cooperative leaves, automatic
                cl.lib().CLSetEngine(engine)
                for _ in range(3):
                    r.execute()
                per_engine[engine] = min(_frame_ms(r) for _ in range(5))
            chosen = int(cl.lib().CLLastEngine())
            ours = per_engine[0]
            row = {"triangles": len(c) // 3, "tree": label, "rays": w * h, "reference_kernel_opencl_ms": ref_ms,
                   "cuda_ms": ours, "speedup": ref_ms / ours, "reference_Mrays_s": w * h / ref_ms / 1e3,
                   "cuda_Mrays_s": w * h / ours / 1e3, "engine_chosen": chosen,
                   "cuda_ms_engine1_lane_per_ray": per_engine[1], "cuda_ms_engine2_cooperative": per_engine[2]}
            print(row, flush=True)
            out["rows"].append(row)
    r.close()
    Path(out_path).write_text(json.dumps(out, indent=1))


def _frame_ms(r):
    r.execute()
    return r.kernel_ms()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "ref_vs_cuda.json")
