import sys, json, numpy as np
sys.path.insert(0,'.')
import clpathtracer_b200 as cl
from clpathtracer_b200 import scenes
from oracle import oracle_py as op
sys.path.insert(0,'tests')
from test_reference_kernel import _cases, _compare
ok, what = op.ref_kernel_available(); print(ok, what)
out={}
for name,(gen,camkw) in _cases().items():
    for (w,h) in [(160,120),(640,480),(1920,1080)]:
        scene = cl.build_kd(*gen()); cam = cl.cam_matrix(cl.make_camera(**camkw), h)
        ref, ms = op.ref_kernel_render(scene, cam, w, h, repeats=3)
        mine = op.render(scene, cam, w, h, mode=0, depth=2)
        s=_compare(ref[...,:3], mine); s["kernel_ms"]=ms; s["build"]=op.ref_kernel_build_options(); out[f"{name}_{w}x{h}"]=s; print(name,w,h,s, flush=True)
# big scenes: reference kernel timing on its own tree (depth 15)
for n in (224, 707):
    scene = cl.build_kd(*scenes.heightfield(n, False)); cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), 1080)
    ref, ms = op.ref_kernel_render(scene, cam, 1920, 1080, repeats=3)
    mine = op.render(scene, cam, 1920, 1080, mode=0, depth=2)
    s=_compare(ref[...,:3], mine); s["kernel_ms"]=ms; s["rays"]=mine["counters"]["rays"]; out[f"hf{n}_1920x1080"]=s; print(n, s, flush=True)
json.dump(out, open('gpurun_out/ref_kernel_stats.json','w'), indent=1)
