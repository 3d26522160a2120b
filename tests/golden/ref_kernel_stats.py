"""Statistics of the oracle restatement against the reference's own kernel.cl ON THE GPU BOX
(oracle/cl_harness.c): as shipped (depth key 0) and with the mirror bounce made runnable
at trace depth 2 and 5.  Writes gpurun_out/ref_kernel_stats.json; the committed copy is
tests/golden/ref_kernel_stats_r02.json."""
import json
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import clpathtracer_b200 as cl  # noqa: E402
from clpathtracer_b200 import scenes  # noqa: E402
from oracle import oracle_py as op  # noqa: E402
from test_reference_kernel import _cases, _compare  # noqa: E402

ok, what = op.ref_kernel_available()
print(ok, what)
out = {"device": what}


def one(key, scene, cam, w, h, depth):
    ref, ms = op.ref_kernel_render(scene, cam, w, h, repeats=3, bounce_depth=depth)
    mine = op.render(scene, cam, w, h, mode=0 if depth == 0 else 1, depth=max(depth, 1))
    s = _compare(ref[..., :3], mine)
    s["kernel_ms"], s["build"], s["rays"] = ms, op.ref_kernel_build_options(), mine["counters"]["rays"]
    out[key] = s
    print(key, s, flush=True)


for name, (gen, camkw) in _cases().items():
    scene = cl.build_kd(*gen())
    for (w, h) in [(160, 120), (640, 480), (1920, 1080)]:
        cam = cl.cam_matrix(cl.make_camera(**camkw), h)
        for depth in (0, 2, 5):
            one(f"{name}_{w}x{h}_d{depth}", scene, cam, w, h, depth)
# big scenes on the reference's own tree (DEPTH 15)
for n in (224, 707):
    scene = cl.build_kd(*scenes.heightfield(n, False))
    cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), 1080)
    for depth in (0, 2, 5):
        one(f"hf{n}_1920x1080_d{depth}", scene, cam, 1920, 1080, depth)
json.dump(out, open("gpurun_out/ref_kernel_stats.json", "w"), indent=1)
