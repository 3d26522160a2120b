"""Generate tests/golden/ref_kernel_golden.npz ON THE GPU BOX.

Runs the reference's own, unmodified src/kernel.cl (linked as a blob into
oracle/_ref/libref_kernel.so by `make -C oracle ref`, see oracle/cl_harness.c)
through the OpenCL ICD of the box's NVIDIA driver, on the deterministic scenes
below, and stores the float4 frames it writes.  The kernel as shipped returns
the first hit's normal colour (src/kernel.cl:395-397), white on a miss; those
frames are stored under the case name.  Under `<case>_d2` and `<case>_d5` it
stores the frames of the reference's mirror bounce (src/kernel.cl:399-417) at
trace depth 2 (the literal at :468) and 5 (the bench depth), executed by the
same vendor compiler from the textual clone oracle/cl_harness.c makes
(bounce_source: the early return removed, the recursion unrolled into named
copies, nothing else).

    gpurun -- 'python tests/golden/make_ref_kernel_golden.py gpurun_out/ref_kernel_golden.npz'
then copy the file to tests/golden/.
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

W, H = 160, 120
CASES = {
    # name: (scene inputs, camera kwargs)
    "hf22_canonical": (lambda: scenes.heightfield(22, False), scenes.CANONICAL_CAMERA),
    "hf22n_canonical": (lambda: scenes.heightfield(22, True), scenes.CANONICAL_CAMERA),
    "hf60_reference": (lambda: scenes.heightfield(60, False), scenes.REFERENCE_CAMERA),
    "cornell": (lambda: scenes.cornell(10)[:3], scenes.CORNELL_CAMERA),
    "soup3000": (lambda: scenes.soup(3000), scenes.CORNELL_CAMERA),
}


def main(out_path):
    ok, what = op.ref_kernel_available()
    if not ok:
        raise SystemExit("no usable OpenCL device: " + what)
    arrays, meta = {}, {"device": what, "width": W, "height": H, "cases": {}}
    for name, (gen, camkw) in CASES.items():
        scene = cl.build_kd(*gen())  # depth 15 / 25 bins: byte-identical to the reference builder
        cam = cl.cam_matrix(cl.make_camera(**camkw), H)
        rgba, ms = op.ref_kernel_render(scene, cam, W, H)
        assert np.all(rgba[..., 3] == 1.0)
        arrays[name] = rgba[..., :3].copy()
        meta["cases"][name] = {"kernel_ms": ms, "hit_fraction": float((rgba[..., :3] != 1.0).any(axis=-1).mean())}
        print(name, meta["cases"][name])
        for d in (2, 5):
            rgba, ms = op.ref_kernel_render(scene, cam, W, H, bounce_depth=d)
            assert np.all(rgba[..., 3] == 1.0)
            arrays[f"{name}_d{d}"] = rgba[..., :3].copy()
            meta["cases"][f"{name}_d{d}"] = {"kernel_ms": ms, "build": op.ref_kernel_build_options()}
            print(name, d, meta["cases"][f"{name}_d{d}"])
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(out_path, **arrays)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).parent / "ref_kernel_golden.npz"))
