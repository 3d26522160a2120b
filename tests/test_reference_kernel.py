"""The reference's own kernel.cl (unmodified, via the box's OpenCL ICD) against
the oracle restatement and against the CUDA path.

This is what pins the oracle: oracle/cl_harness.c runs the reference's device
code as shipped (first-hit normal colour) on the same 68-byte node arrays.  The
vendor's OpenCL compiler contracts FMAs and uses its own normalize()/divide, so
agreement is not bit-level; the bar is SURVEY.md section 8d (ii)/(iii):
pixels that land on a different triangle (or flip hit/miss) are classified --
exact edge grazes (a tie between the two triangles sharing an edge) are counted
separately, everything else must stay <= 1e-4 of the frame -- and the mean abs
error must be < 1e-3.  Where both resolve the same triangle the colours agree to
a few ulp.  NVIDIA's OpenCL compiler rejects the unmodified source (a __global
pointer passed to an unqualified parameter); oracle/cl_harness.c documents the
one-declaration address-space patch it applies in memory, nothing else.

GPU half (-m gpu): live run on the box.  CPU half: the frames such a run
produced, committed as tests/golden/ref_kernel_golden.npz.
"""
import json
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).parent / "golden" / "ref_kernel_golden.npz"
MISMATCH_BUDGET = 1e-4   # unclassified mismatches allowed per traced RAY (for the as-shipped mode: per pixel)
MAE_TOL = 1e-3           # north-star radiance tolerance
SAME_PATH_TOL = 1e-6     # colour agreement where the same triangles are hit (a few ulp of 0.5..1, blended)
EDGE_EPS = 2e-6          # a hit this close to a triangle edge (in barycentrics) is an edge graze


def _compare(ref_rgb, mine):
    """mine: dict with rgba/path_edge/counters from the oracle (or rgba from CUDA + the oracle's AOVs).
    A pixel 'mismatches' when the reference kernel's colour differs by more than SAME_PATH_TOL:
    it resolved a different triangle (or flipped hit/miss) somewhere along the pixel's path.
    Mismatches whose path touches a triangle edge (u, v or 1-u-v within EDGE_EPS of 0 at the
    primary hit or any bounce) are exact ties between the two triangles sharing that edge --
    which one wins is decided by the last bit of t, i.e. by FMA contraction in the vendor
    compiler -- and are classified as edge grazes.  Everything else
    counts against the budget, which is per RAY: the as-shipped mode traces one ray per pixel
    (the survey's 1e-4 of pixels); with bounces every ray of the path has the same chance of
    an unclassifiable last-bit difference, and one ulp in a normal moves the NEXT hit."""
    mine_rgb = mine["rgba"][..., :3]
    d = np.abs(ref_rgb.astype(np.float64) - mine_rgb.astype(np.float64)).max(axis=-1)
    different = d > SAME_PATH_TOL
    on_edge = mine["path_edge"] <= EDGE_EPS
    unclassified = different & ~on_edge
    rays = max(int(mine["counters"]["rays"]), d.size)
    return {"mismatch_fraction": float(different.mean()), "edge_graze_pixels": int((different & on_edge).sum()),
            "unclassified_pixels": int(unclassified.sum()), "rays": rays,
            "unclassified_per_ray": float(unclassified.sum() / rays),
            "visible_mismatch_fraction": float((d > 1e-4).mean()),
            "mae": float(np.abs(ref_rgb - mine_rgb).mean()),
            "max_same_path_err": float(d[~different].max()) if (~different).any() else 0.0}


def _check(s):
    assert s["unclassified_per_ray"] <= MISMATCH_BUDGET, s
    assert s["mismatch_fraction"] <= 20 * MISMATCH_BUDGET, s
    assert s["mae"] < MAE_TOL, s


def _cases():
    from clpathtracer_b200 import scenes

    return {
        "hf22_canonical": (lambda: scenes.heightfield(22, False), scenes.CANONICAL_CAMERA),
        "hf22n_canonical": (lambda: scenes.heightfield(22, True), scenes.CANONICAL_CAMERA),
        "hf60_reference": (lambda: scenes.heightfield(60, False), scenes.REFERENCE_CAMERA),
        "cornell": (lambda: scenes.cornell(10)[:3], scenes.CORNELL_CAMERA),
        "soup3000": (lambda: scenes.soup(3000), scenes.CORNELL_CAMERA),
    }


def _oracle_frame(oracle, scene, cam, w, h, bounce_depth):
    """bounce_depth 0: as shipped (mode A).  D >= 1: the mirror bounce at trace depth D (mode B)."""
    if bounce_depth == 0:
        return oracle.render(scene, cam, w, h, mode=0, depth=2)
    return oracle.render(scene, cam, w, h, mode=1, depth=bounce_depth)


@pytest.mark.skipif(not GOLDEN.exists(), reason="no reference-kernel golden committed yet")
@pytest.mark.parametrize("bounce_depth", [0, 2, 5])
@pytest.mark.parametrize("name", ["hf22_canonical", "hf22n_canonical", "hf60_reference", "cornell", "soup3000"])
def test_oracle_vs_reference_kernel_golden(clpt, oracle, name, bounce_depth):
    """Depth 0 = the kernel as shipped.  Depth 2 (the literal at src/kernel.cl:468) and 5
    (the bench depth) = the reference's mirror bounce (:399-417) as executed by the vendor
    compiler from the textual clone of oracle/cl_harness.c -- the pin of mode B."""
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    w, h = meta["width"], meta["height"]
    key = name if bounce_depth == 0 else f"{name}_d{bounce_depth}"
    if key not in z.files:
        pytest.skip(f"golden has no {key} (regenerate with tests/golden/make_ref_kernel_golden.py)")
    gen, camkw = _cases()[name]
    scene = clpt.build_kd(*gen())
    cam = clpt.cam_matrix(clpt.make_camera(**camkw), h)
    mine = _oracle_frame(oracle, scene, cam, w, h, bounce_depth)
    s = _compare(z[key], mine)
    _check(s)


def test_bounce_rewrite_is_textual(oracle):
    """The source the vendor compiler gets for a bounce render differs from the reference's
    kernel.cl only by whole cloned copies of trace_ray, the removed early-return statement,
    the renamed callee and the depth literal (oracle/cl_harness.c: bounce_source)."""
    import difflib

    ref_src = Path("/root/reference/src/kernel.cl")
    if not ref_src.exists() or not oracle.REF_KERNEL_LIB.exists():
        pytest.skip("needs /root/reference and oracle/_ref (build container only)")
    orig = ref_src.read_text()
    lines = orig.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("trace_ray(Ray r,")) - 1
    end = next(i for i, l in enumerate(lines) if l.startswith("kernel void"))
    fn = lines[start:end]
    for depth in (2, 5):
        got = oracle.ref_kernel_bounce_source(depth).splitlines()
        assert got[:start] == lines[:start]  # everything before trace_ray is untouched
        assert len(got) == len(lines) + depth * len(fn) - depth  # `depth` clones; each live copy lost one line
        changed = [l for l in difflib.ndiff(fn, got[start:start + len(fn)]) if l[0] in "+-"]
        # deepest clone = the function as shipped, renamed (definition + its dead self-call)
        assert sorted(c[2:].strip() for c in changed) == sorted(
            ["trace_ray(Ray r,", f"trace_ray_{depth}(Ray r,", "return trace_ray(newRay,", f"return trace_ray_{depth}(newRay,"])
        tail = got[start + depth * len(fn) - (depth - 1):]  # the outermost copy + render
        diff = [l for l in difflib.ndiff(lines[start:], tail) if l[0] in "+-"]
        removed = [l[2:].strip() for l in diff if l[0] == "-"]
        added = [l[2:].strip() for l in diff if l[0] == "+"]
        assert removed[0].startswith("return convert_color((normal + 1) / 2)/*")
        assert removed[1].endswith("*/;")
        assert removed[2] == "return trace_ray(newRay,"
        assert added[:2] == ["", "return trace_ray_1(newRay,"]
        assert (removed[3:], added[2:]) == ((["2,"], [f"{depth},"]) if depth != 2 else ([], []))


@pytest.mark.gpu
@pytest.mark.parametrize("bounce_depth", [0, 2, 5])
@pytest.mark.parametrize("name,w,h", [("hf22_canonical", 640, 480), ("cornell", 640, 480),
                                      ("hf60_reference", 480, 270), ("soup3000", 320, 320)])
def test_reference_kernel_live(clpt, oracle, renderer, name, w, h, bounce_depth):
    ok, what = oracle.ref_kernel_available()
    if not ok:
        pytest.skip("reference kernel cannot run here: " + what)
    gen, camkw = _cases()[name]
    scene = clpt.build_kd(*gen())
    cam = clpt.cam_matrix(clpt.make_camera(**camkw), h)
    ref_rgba, _ = oracle.ref_kernel_render(scene, cam, w, h, bounce_depth=bounce_depth)
    assert np.all(ref_rgba[..., 3] == 1.0)
    ref = _oracle_frame(oracle, scene, cam, w, h, bounce_depth)
    want = ref["rgba"]
    renderer.set_meshes(scene)
    renderer.set_camera_matrix(cam)
    renderer.set_params(mode=0 if bounce_depth == 0 else 1, depth=2 if bounce_depth == 0 else bounce_depth)
    renderer.create_image(w, h)
    renderer.execute()
    got = renderer.read_image()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # CUDA == oracle, bit for bit
    s = _compare(ref_rgba[..., :3], dict(ref, rgba=got))                # CUDA vs the reference kernel itself
    _check(s)
