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
MISMATCH_BUDGET = 1e-4   # fraction of pixels allowed to resolve a different triangle away from any edge
MAE_TOL = 1e-3           # north-star radiance tolerance
SAME_TRI_TOL = 2e-6      # colour agreement where the same triangle is hit (~ a few ulp of 0.5..1)


EDGE_EPS = 2e-6          # a hit this close to a triangle edge (in barycentrics) is an edge graze


def _compare(ref_rgb, mine):
    """mine: dict with rgba/prim/uv from the oracle (or rgba from CUDA + the oracle's AOVs).
    A pixel 'mismatches' when the reference kernel resolved a different triangle or
    flipped hit/miss.  Mismatches whose hit lies on a triangle edge (u, v or 1-u-v
    within EDGE_EPS of 0) are exact ties between the two triangles sharing that edge
    -- which one wins is decided by the last bit of t, i.e. by FMA contraction in the
    vendor compiler -- and are classified as edge grazes.
    Everything else counts against the budget."""
    mine_rgb = mine["rgba"][..., :3]
    d = np.abs(ref_rgb.astype(np.float64) - mine_rgb.astype(np.float64)).max(axis=-1)
    different = d > 1e-4
    u, v = mine["uv"][..., 0].astype(np.float64), mine["uv"][..., 1].astype(np.float64)
    on_edge = (mine["prim"] >= 0) & (np.minimum(np.minimum(u, v), np.abs(1.0 - u - v)) <= EDGE_EPS)
    unclassified = different & ~on_edge
    return {"mismatch_fraction": float(different.mean()), "edge_graze_pixels": int((different & on_edge).sum()),
            "unclassified_fraction": float(unclassified.mean()), "unclassified_pixels": int(unclassified.sum()),
            "mae": float(np.abs(ref_rgb - mine_rgb).mean()),
            "max_same_tri_err": float(d[~different].max()) if (~different).any() else 0.0}


def _cases():
    from clpathtracer_b200 import scenes

    return {
        "hf22_canonical": (lambda: scenes.heightfield(22, False), scenes.CANONICAL_CAMERA),
        "hf22n_canonical": (lambda: scenes.heightfield(22, True), scenes.CANONICAL_CAMERA),
        "hf60_reference": (lambda: scenes.heightfield(60, False), scenes.REFERENCE_CAMERA),
        "cornell": (lambda: scenes.cornell(10)[:3], scenes.CORNELL_CAMERA),
        "soup3000": (lambda: scenes.soup(3000), scenes.CORNELL_CAMERA),
    }


@pytest.mark.skipif(not GOLDEN.exists(), reason="no reference-kernel golden committed yet")
@pytest.mark.parametrize("name", ["hf22_canonical", "hf22n_canonical", "hf60_reference", "cornell", "soup3000"])
def test_oracle_vs_reference_kernel_golden(clpt, oracle, name):
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    w, h = meta["width"], meta["height"]
    gen, camkw = _cases()[name]
    scene = clpt.build_kd(*gen())
    cam = clpt.cam_matrix(clpt.make_camera(**camkw), h)
    mine = oracle.render(scene, cam, w, h, mode=0, depth=2)
    s = _compare(z[name], mine)
    assert s["unclassified_fraction"] <= MISMATCH_BUDGET, s
    assert s["mismatch_fraction"] <= 20 * MISMATCH_BUDGET, s
    assert s["mae"] < MAE_TOL, s
    assert s["max_same_tri_err"] <= SAME_TRI_TOL, s


@pytest.mark.gpu
@pytest.mark.parametrize("name,w,h", [("hf22_canonical", 640, 480), ("cornell", 640, 480),
                                      ("hf60_reference", 480, 270), ("soup3000", 320, 320)])
def test_reference_kernel_live(clpt, oracle, renderer, name, w, h):
    ok, what = oracle.ref_kernel_available()
    if not ok:
        pytest.skip("reference kernel cannot run here: " + what)
    gen, camkw = _cases()[name]
    scene = clpt.build_kd(*gen())
    cam = clpt.cam_matrix(clpt.make_camera(**camkw), h)
    ref_rgba, _ = oracle.ref_kernel_render(scene, cam, w, h)
    assert np.all(ref_rgba[..., 3] == 1.0)
    ref = oracle.render(scene, cam, w, h, mode=0, depth=2)
    want = ref["rgba"]
    renderer.set_meshes(scene)
    renderer.set_camera_matrix(cam)
    renderer.set_params(mode=0, depth=2)
    renderer.create_image(w, h)
    renderer.execute()
    got = renderer.read_image()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # CUDA == oracle, bit for bit
    s = _compare(ref_rgba[..., :3], dict(ref, rgba=got))                # CUDA vs the reference kernel itself
    assert s["unclassified_fraction"] <= MISMATCH_BUDGET, s
    assert s["mismatch_fraction"] <= 20 * MISMATCH_BUDGET, s
    assert s["mae"] < MAE_TOL, s
    assert s["max_same_tri_err"] <= SAME_TRI_TOL, s
