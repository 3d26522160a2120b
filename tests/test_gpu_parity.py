"""Parity of the CUDA render path (through the C ABI) with the oracle.

Every test drives libclpt.so through the CLState entry points and compares with
oracle/oracle_kernel.c on the same scene, camera and seed.  Bars (SURVEY.md
section 8d): first-hit primitive ids, t, (u,v) and colours BIT-EXACT for the
deterministic modes (both sides evaluate the same fp32 expressions in the same
order without FMA contraction); the stochastic extension is compared with the
same Philox streams and must also be bit-exact; the north-star's looser bound
(mean abs error < 1e-3 at 1 spp) is asserted as well, with the tolerance
written here.
"""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MAE_TOL = 1e-3  # north-star tolerance for accumulated radiance at 1 spp


def _cam(clpt, which, height):
    from clpathtracer_b200 import scenes

    kw = {"canonical": scenes.CANONICAL_CAMERA, "reference": scenes.REFERENCE_CAMERA,
          "cornell": scenes.CORNELL_CAMERA}[which]
    return clpt.cam_matrix(clpt.make_camera(**kw), height)


def _render_gpu(r, scene, cam, w, h, aov=True, **params):
    r.set_meshes(scene)
    r.set_camera_matrix(cam)
    r.set_params(**params)
    r.create_image(w, h, aov=aov)
    r.execute()
    img = r.read_image()
    return (img,) + (r.read_aov() if aov else ())


def _assert_bit_equal(a, b, what):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    same = a.view(np.uint32) == b.view(np.uint32)
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {same.size} words differ"


CASES = [
    # scene, camera, w, h, mode, depth
    ("hf22n", "canonical", 640, 480, 0, 2),   # BASELINE config 1 size, as shipped
    ("hf22n", "canonical", 640, 480, 1, 2),   # primary + 1 mirror bounce
    ("hf22", "canonical", 320, 240, 0, 2),    # flat normals (no vn)
    ("hf22", "canonical", 320, 240, 1, 5),
    ("hf22n", "reference", 333, 197, 1, 3),   # odd sizes, camera inside the scene box
    ("cornell", "cornell", 640, 480, 0, 2),
    ("cornell", "cornell", 640, 480, 1, 5),
    ("soup3000", "cornell", 256, 256, 1, 4),  # incoherent
    ("hf224", "canonical", 480, 270, 1, 2),   # 100k triangles
]


@pytest.mark.parametrize("name,camera,w,h,mode,depth", CASES)
def test_deterministic_modes_bit_exact(clpt, oracle, renderer, scene_cache, name, camera, w, h, mode, depth):
    scene, _ = scene_cache(name)
    cam = _cam(clpt, camera, h)
    img, prim, t, uv = _render_gpu(renderer, scene, cam, w, h, mode=mode, depth=depth)
    ref = oracle.render(scene, cam, w, h, mode=mode, depth=depth)
    assert np.array_equal(prim, ref["prim"]), f"{np.count_nonzero(prim != ref['prim'])} primitive ids differ"
    _assert_bit_equal(t, ref["t"], "t")
    _assert_bit_equal(uv, ref["uv"], "uv")
    _assert_bit_equal(img, ref["rgba"], "rgba")
    assert np.abs(img - ref["rgba"]).mean() < MAE_TOL
    assert (prim >= 0).mean() > (0.01 if name.startswith('soup') else 0.05)  # the camera sees the scene


@pytest.mark.parametrize("name,camera,w,h,mode,depth,sah", [("hf224", "canonical", 480, 270, 1, 5, True),
                                                             ("cornell", "cornell", 320, 240, 1, 4, True),
                                                             ("soup3000", "cornell", 256, 256, 1, 4, True),
                                                             ("hf22n", "reference", 333, 197, 0, 2, True),
                                                             ("hf224", "canonical", 480, 270, 1, 5, "exact"),
                                                             ("soup3000", "cornell", 256, 256, 1, 4, "exact")])
def test_sah_trees_bit_exact(clpt, oracle, renderer, scene_cache, name, camera, w, h, mode, depth, sah):
    """Trees from the SAH builder (extension; binned and exact-sweep) go through the same traversal."""
    scene, _ = scene_cache(name, sah=sah)
    cam = _cam(clpt, camera, h)
    img, prim, t, uv = _render_gpu(renderer, scene, cam, w, h, mode=mode, depth=depth)
    ref = oracle.render(scene, cam, w, h, mode=mode, depth=depth)
    assert np.array_equal(prim, ref["prim"])
    _assert_bit_equal(t, ref["t"], "t")
    _assert_bit_equal(uv, ref["uv"], "uv")
    _assert_bit_equal(img, ref["rgba"], "rgba")


COOP_CASES = [
    # scene, tree, camera, w, h, mode, depth, spp, flags
    ("hf224", False, "canonical", 480, 270, 0, 2, 1, 0),      # reference tree, ~21 triangles per leaf, as shipped
    ("hf224", False, "canonical", 480, 270, 1, 5, 1, 0),      # mirror bounces
    ("hf224", False, "canonical", 333, 197, 1, 3, 6, 1),      # spp not a power of two, jitter, odd size
    ("hf22n", 6, "canonical", 320, 240, 1, 4, 4, 1),          # a very shallow tree: 30+ triangles per leaf, vn shading
    ("soup3000", 5, "cornell", 256, 256, 1, 4, 2, 1),         # incoherent, fat leaves
    ("cornell", False, "cornell", 320, 240, 2, 4, 8, 1),      # path mode through the cooperative loop
    ("hf22n", False, "canonical", 200, 150, 1, 1, 1, 0),      # thin leaves only: the per-lane branch of engine 2
    ("hf22n", False, "canonical", 64, 48, 1, 0, 1, 0),        # depth 0: nothing is traced, AOVs say miss
]


@pytest.mark.parametrize("name,tree,camera,w,h,mode,depth,spp,flags", COOP_CASES)
def test_fat_leaf_engine_bit_exact(clpt, oracle, renderer, scene_cache, name, tree, camera, w, h, mode, depth, spp,
                                      flags):
    """Engine 2 (the kernel compiled for fewer resident blocks, chosen for trees with fat
    leaves) produces the same bits as the oracle and as engine 1, work counters included."""
    from clpathtracer_b200 import scenes

    if isinstance(tree, bool):
        scene, _ = scene_cache(name, sah=tree)
    else:  # the reference heuristic stopped at a small depth: fat leaves on a small scene
        gen = {"hf22n": lambda: scenes.heightfield(22, True), "soup3000": lambda: scenes.soup(3000)}[name]
        scene = clpt.build_kd(*gen(), depth=tree)
    assert scene.stats()["max_leaf_tris"] >= 8 or name in ("hf22n", "cornell")
    cam = _cam(clpt, camera, h)
    L = clpt.lib()
    kw = dict(mode=mode, depth=depth, spp=spp, seed=5)
    okw = {}
    if mode == 2:  # one diffuse material with a little emission (set_meshes keeps the material table)
        grey = np.array([[0.7, 0.6, 0.5, 0, 0.1, 0.1, 0.2, 0]], dtype=np.float32)
        renderer.set_materials(grey, None)
        okw["materials"] = grey
    try:
        L.CLSetEngine(2)
        img, prim, t, uv = _render_gpu(renderer, scene, cam, w, h, flags=flags | clpt.FLAG_COUNTERS, **kw)
        assert L.CLLastEngine() == 2
        got_counters = renderer.counters()
        plain = _render_gpu(renderer, scene, cam, w, h, flags=flags, **kw)[0]
        L.CLSetEngine(1)
        lane = _render_gpu(renderer, scene, cam, w, h, flags=flags, **kw)[0]
        assert L.CLLastEngine() == 1
    finally:
        L.CLSetEngine(0)
    ref = oracle.render(scene, cam, w, h, flags=flags, **kw, **okw)
    assert np.array_equal(prim, ref["prim"])
    _assert_bit_equal(t, ref["t"], "t")
    _assert_bit_equal(uv, ref["uv"], "uv")
    _assert_bit_equal(img, ref["rgba"], "rgba")
    _assert_bit_equal(plain, ref["rgba"], "rgba (uninstrumented kernel)")
    _assert_bit_equal(img, lane, "engine 2 vs engine 1")
    assert got_counters == ref["counters"]


def test_engine_is_chosen_per_tree(clpt, renderer, scene_cache):
    """Automatic engine: the fat-leaf variant for big trees whose triangles sit in fat leaves (the
    reference builder from a few hundred thousand triangles up), the full-occupancy kernel for
    small ones and for SAH trees."""
    L = clpt.lib()
    L.CLSetEngine(0)
    cam = _cam(clpt, "canonical", 48)
    from clpathtracer_b200 import scenes

    big = clpt.build_kd(*scenes.heightfield(520, False))  # 540,800 triangles, DEPTH 15: ~40 per leaf
    for scene, want in ((big, 2), (scene_cache("hf224", sah=False)[0], 1), (scene_cache("hf224", sah=True)[0], 1)):
        _render_gpu(renderer, scene, cam, 64, 48, aov=False, mode=0, depth=2)
        assert L.CLLastEngine() == want


@pytest.mark.parametrize("depth_tree", [8, 20, 24])
def test_other_tree_depths(clpt, oracle, renderer, depth_tree):
    """The traversal is tree-agnostic: shallow and deep trees of the same mesh."""
    from clpathtracer_b200 import scenes

    v, c, n = scenes.heightfield(100, False)
    scene = clpt.build_kd(v, c, n, depth=depth_tree)
    cam = _cam(clpt, "canonical", 270)
    img, prim, t, uv = _render_gpu(renderer, scene, cam, 480, 270, mode=1, depth=3)
    ref = oracle.render(scene, cam, 480, 270, mode=1, depth=3)
    assert np.array_equal(prim, ref["prim"])
    _assert_bit_equal(t, ref["t"], "t")
    _assert_bit_equal(img, ref["rgba"], "rgba")


def test_counters_match_oracle(clpt, oracle, renderer, scene_cache):
    """Same leaves, same triangles, same order: the work counters are equal."""
    scene, _ = scene_cache("hf22n")
    cam = _cam(clpt, "canonical", 480)
    _render_gpu(renderer, scene, cam, 640, 480, mode=1, depth=2, flags=clpt.FLAG_COUNTERS)
    got = renderer.counters()
    want = oracle.render(scene, cam, 640, 480, mode=1, depth=2)["counters"]
    assert got == want
    # SURVEY.md section 6 figures for this scene/camera, as shipped
    _render_gpu(renderer, scene, cam, 640, 480, mode=0, depth=2, flags=clpt.FLAG_COUNTERS)
    got = renderer.counters()
    assert got["rays"] == 307200
    assert round(got["splits"] / got["rays"], 2) == 12.55 and round(got["tris"] / got["rays"], 2) == 5.17


def test_jittered_multisample(clpt, oracle, renderer, scene_cache):
    scene, _ = scene_cache("hf22n")
    cam = _cam(clpt, "canonical", 240)
    kw = dict(mode=1, depth=3, spp=8, seed=1234, flags=clpt.FLAG_JITTER)
    img, prim, t, uv = _render_gpu(renderer, scene, cam, 320, 240, **kw)
    ref = oracle.render(scene, cam, 320, 240, **kw)
    assert np.array_equal(prim, ref["prim"])
    _assert_bit_equal(img, ref["rgba"], "rgba")
    # a different seed gives a different image; the same seed the same image
    img2 = _render_gpu(renderer, scene, cam, 320, 240, **dict(kw, seed=99))[0]
    assert not np.array_equal(img, img2)
    img3 = _render_gpu(renderer, scene, cam, 320, 240, **kw)[0]
    assert np.array_equal(img, img3)
    # spp == 1 without jitter is the reference's own ray generation
    a = _render_gpu(renderer, scene, cam, 320, 240, mode=1, depth=3, spp=1, seed=5)[0]
    b = _render_gpu(renderer, scene, cam, 320, 240, mode=1, depth=3, spp=1, seed=77)[0]
    assert np.array_equal(a, b)


@pytest.mark.parametrize("spp,engine,mode,warps", [(64, 1, 1, 0), (130, 1, 1, 0), (300, 1, 1, 0), (64, 2, 1, 0),
                                                   (33, 1, 1, 0), (64, 1, 1, 1), (130, 1, 2, 0), (64, 1, 0, 0),
                                                   (300, 1, 2, 2)])
def test_samples_spread_over_warps(clpt, oracle, renderer, scene_cache, monkeypatch, spp, engine, mode, warps):
    """At >= 64 spp (small frames / sharded frames) the samples of a pixel are traced by 2, 4 or 8
    warps side by side and summed by the first of them in ascending sample order: the image does
    not depend on the spread (64 -> 2 warps, one round; 130 -> 4 warps, a full round and one of
    2 samples; 300 -> 8 warps; 33 -> one warp, two rounds; warps != 0 forces a spread)."""
    scene, extra = scene_cache("cornell" if mode == 2 else "hf22n")
    w, h = 61, 37
    cam = _cam(clpt, "cornell" if mode == 2 else "canonical", h)
    kw = dict(mode=mode, depth=4, spp=spp, seed=3, flags=clpt.FLAG_JITTER)
    okw = {}
    if warps:
        monkeypatch.setenv("CLPT_WARPS_PER_PIXEL", str(warps))
    L = clpt.lib()
    try:
        L.CLSetEngine(engine)
        if mode == 2:
            from clpathtracer_b200 import scenes

            renderer.set_meshes(scene)
            renderer.set_materials(scenes.CORNELL_MATERIALS, extra["tri_material"])
            okw = dict(materials=scenes.CORNELL_MATERIALS, tri_material=extra["tri_material"])
            renderer.set_camera_matrix(cam)
            renderer.set_params(**kw)
            renderer.create_image(w, h, aov=True)
            renderer.execute()
            img = renderer.read_image()
            prim, t, uv = renderer.read_aov()
        else:
            img, prim, t, uv = _render_gpu(renderer, scene, cam, w, h, **kw)
    finally:
        L.CLSetEngine(0)
    ref = oracle.render(scene, cam, w, h, **kw, **okw)
    assert np.array_equal(prim, ref["prim"])
    _assert_bit_equal(img, ref["rgba"], f"{spp} spp")


def test_path_mode_extension(clpt, oracle, renderer, scene_cache):
    """Mode C (no reference behaviour): identical Philox streams -> identical images."""
    from clpathtracer_b200 import scenes

    scene, extra = scene_cache("cornell")
    cam = _cam(clpt, "cornell", 240)
    renderer.set_meshes(scene)
    renderer.set_materials(scenes.CORNELL_MATERIALS, extra["tri_material"])
    renderer.set_camera_matrix(cam)
    kw = dict(mode=2, depth=4, spp=4, seed=7, flags=clpt.FLAG_JITTER)
    renderer.set_params(**kw)
    renderer.create_image(320, 240, aov=True)
    renderer.execute()
    img = renderer.read_image()
    ref = oracle.render(scene, cam, 320, 240, materials=scenes.CORNELL_MATERIALS,
                        tri_material=extra["tri_material"], **kw)
    _assert_bit_equal(img, ref["rgba"], "rgba")
    assert np.abs(img - ref["rgba"]).mean() < MAE_TOL
    assert img[..., :3].max() > 1.0  # the light is visible
    # degrade: mirror materials at 1 spp without jitter reproduce the mirror geometry of mode B
    mirror = np.array([[1, 1, 1, 0, 0, 0, 0, 0]], dtype=np.float32)
    mirror.view(np.int32)[0, 3] = 1
    renderer.set_materials(mirror, None)
    renderer.set_params(mode=2, depth=3, spp=1, seed=0, flags=0)
    renderer.execute()
    ref = oracle.render(scene, cam, 320, 240, mode=2, depth=3, materials=mirror)
    _assert_bit_equal(renderer.read_image(), ref["rgba"], "mirror path rgba")


def test_progressive_accumulation(clpt, oracle, renderer, scene_cache):
    """Progressive frames add their samples to 2^-32 fixed-point sums (order-free integer
    additions); the read-back is the mean.  Same bits as the oracle's, whatever the grouping of
    the samples into frames."""
    scene, _ = scene_cache("hf22n")
    cam = _cam(clpt, "canonical", 120)
    flags = clpt.FLAG_JITTER | clpt.FLAG_ACCUMULATE
    renderer.set_meshes(scene)
    renderer.set_camera_matrix(cam)
    renderer.set_params(mode=1, depth=2, spp=2, seed=3, flags=flags)
    renderer.create_image(160, 120)
    acc = oracle.new_accumulator(160, 120)
    for frame in range(3):
        renderer.execute()
        want = oracle.render(scene, cam, 160, 120, mode=1, depth=2, spp=2, seed=3, flags=flags, sample_base=2 * frame,
                             accumulate_into=acc, aov=False)["rgba"]
    _assert_bit_equal(renderer.read_image(), want, "accumulated rgba")
    # the same six samples in one oracle call, and as 3 + 3: the sums do not depend on the grouping
    once = oracle.render(scene, cam, 160, 120, mode=1, depth=2, spp=6, seed=3, flags=flags, aov=False)
    assert np.array_equal(once["accum"], acc)
    _assert_bit_equal(once["rgba"], want, "grouping")
    assert np.abs(want - oracle.render(scene, cam, 160, 120, mode=1, depth=2, spp=6, seed=3, flags=clpt.FLAG_JITTER,
                                       aov=False)["rgba"]).max() < 1e-6  # (the float mean agrees to rounding)
    clpt.lib().CLResetAccumulation()
    renderer.execute()
    one = renderer.read_image()
    ref = oracle.render(scene, cam, 160, 120, mode=1, depth=2, spp=2, seed=3, flags=flags, aov=False)
    _assert_bit_equal(one, ref["rgba"], "after reset")
    renderer.set_params()


@pytest.mark.parametrize("nranks,tile_rows", [(2, 8), (3, 16), (8, 8)])
def test_row_tile_sharding(clpt, oracle, renderer, scene_cache, nranks, tile_rows):
    """Each rank's rows, rendered separately, are exactly those rows of the full frame."""
    scene, _ = scene_cache("hf22n")
    w, h = 200, 150
    cam = _cam(clpt, "canonical", h)
    full = oracle.render(scene, cam, w, h, mode=1, depth=3)["rgba"]
    L = clpt.lib()
    seen = np.zeros(h, dtype=int)
    try:
        for rank in range(nranks):
            renderer.set_meshes(scene)
            renderer.set_camera_matrix(cam)
            renderer.set_params(mode=1, depth=3)
            renderer.create_image(w, h)
            L.CLSetTileShard(rank, nranks, tile_rows)
            renderer.execute()
            img = renderer.read_image()
            rows = np.array([y for y in range(h) if (y // tile_rows) % nranks == rank], dtype=int)
            seen[rows] += 1
            _assert_bit_equal(img[rows], full[rows], f"rank {rank} rows")
            others = np.setdiff1d(np.arange(h), rows)
            assert not img[others].any()  # untouched without a communicator
    finally:
        L.CLSetTileShard(0, 1, 8)
    assert (seen == 1).all()


def test_owned_meshes_and_object_slot(clpt, oracle, renderer, scene_cache):
    """The reference's own CLSetMeshes(kd*) ownership path and the no-op object slot."""
    scene, _ = scene_cache("hf22")
    cam = _cam(clpt, "canonical", 96)
    renderer.set_meshes_owned(scene)
    L = clpt.lib()
    L.CLSetObjects(None, 0)
    spheres = np.zeros(48, dtype=np.uint8)  # two 24-byte Objects
    L.CLSetObjects(spheres.ctypes.data, spheres.nbytes)
    empty_models = L.new_list(0)
    L.CLSetMeshes(empty_models)  # empty vector: no-op (src/CLState.c:126-129)
    L.delete_list(empty_models)
    renderer.set_camera_matrix(cam)
    renderer.set_params(mode=0, depth=2)
    renderer.create_image(128, 96)
    renderer.execute()
    ref = oracle.render(scene, cam, 128, 96, mode=0, depth=2)
    _assert_bit_equal(renderer.read_image(), ref["rgba"], "rgba")
    assert L.CLLastLaunchCount() == 1 and L.CLLastKernelMs() > 0


def test_scene_reupload(clpt, oracle, renderer):
    """The animated path: CLSetMeshes again with moved geometry (exact-size
    re-creation of every device buffer, src/CLState.c:92-102) between frames."""
    from clpathtracer_b200 import scenes

    tv, tc, _ = scenes.heightfield(30, False)
    cam = _cam(clpt, "canonical", 120)
    renderer.create_image(160, 120, aov=True)
    renderer.set_params(mode=1, depth=2)
    renderer.set_camera_matrix(cam)
    prev = None
    for k, builder in enumerate([clpt.build_kd, clpt.build_kd_sah, clpt.build_kd_sah]):
        verts = tv.copy()
        verts[:, 1] += np.float32(0.05 * k)  # the whole mesh moves up
        scene = builder(verts, tc, None)
        renderer.set_meshes(scene)
        renderer.execute()
        img = renderer.read_image()
        ref = oracle.render(scene, cam, 160, 120, mode=1, depth=2)
        _assert_bit_equal(img, ref["rgba"], f"frame {k}")
        assert prev is None or not np.array_equal(prev, img)
        prev = img


def test_clhandler_layer_and_leaf_cap(clpt, oracle, renderer, scene_cache):
    """The CLHandler.h wrappers (the reference's runtime layer, include/CLHandler.h:6-25)
    drive the same launch as CLExecute; the rope-hop cap matches the oracle's."""
    scene, _ = scene_cache("hf22n")
    cam = _cam(clpt, "canonical", 96)
    L = clpt.lib()
    renderer.set_meshes(scene)
    renderer.set_camera_matrix(cam)
    renderer.set_params(mode=1, depth=2)
    renderer.create_image(128, 96)
    for fn, res, args in [("CLGetPlatform", C.c_void_p, []), ("CLGetDevice", C.c_void_p, [C.c_void_p]),
                          ("CLCreateKernel", C.c_void_p, [C.c_char_p, C.c_void_p]),
                          ("CLCreateBuffer", C.c_void_p, [C.c_void_p, C.c_size_t]),
                          ("CLReleaseBuffer", None, [C.c_void_p]),
                          ("CLEnqueueKernel", None, [C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
                          ("err_string", C.c_char_p, [C.c_int])]:
        getattr(L, fn).restype, getattr(L, fn).argtypes = res, args
    plat = L.CLGetPlatform()
    assert L.CLGetDevice(plat)
    kern = L.CLCreateKernel(b"render", None)
    buf = L.CLCreateBuffer(None, 4096)
    assert buf
    L.CLReleaseBuffer(buf)
    size = (C.c_size_t * 2)(128, 96)
    L.CLEnqueueKernel(2, size, None, None, kern)  # global = {w, h}, local = NULL (src/CLState.c:209-211)
    ref = oracle.render(scene, cam, 128, 96, mode=1, depth=2)
    _assert_bit_equal(renderer.read_image(), ref["rgba"], "CLEnqueueKernel frame")
    assert L.err_string(2) == b"cudaErrorMemoryAllocation" and L.err_string(0) == b"cudaSuccess"
    try:
        L.CLSetMaxLeafVisits(2)
        renderer.set_params(mode=1, depth=2, flags=clpt.FLAG_COUNTERS)
        renderer.execute()
        capped = oracle.render(scene, cam, 128, 96, mode=1, depth=2, max_leaf_visits=2)
        _assert_bit_equal(renderer.read_image(), capped["rgba"], "capped frame")
        assert renderer.counters() == capped["counters"] and capped["counters"]["capped"] > 0
        # the kernel counts the rope-hop budget DOWN; the oracle counts hops up against the cap,
        # with and without the instrumented twin (caps below 1 are refused by CLSetMaxLeafVisits)
        for cap in (1, 3, 7):
            L.CLSetMaxLeafVisits(cap)
            capped = oracle.render(scene, cam, 128, 96, mode=1, depth=3, max_leaf_visits=cap)
            for flags in (0, clpt.FLAG_COUNTERS):
                renderer.set_params(mode=1, depth=3, flags=flags)
                renderer.execute()
                _assert_bit_equal(renderer.read_image(), capped["rgba"], f"cap {cap} flags {flags}")
                if flags:
                    assert renderer.counters() == capped["counters"]
    finally:
        L.CLSetMaxLeafVisits(4096)
        renderer.set_params()


@pytest.mark.parametrize("sah", [False, True])
def test_degenerate_rays_and_random_views(clpt, oracle, renderer, scene_cache, sah):
    """Axis-aligned views (direction components exactly 0 -> infinite reciprocals,
    0*inf = NaN in the slab tests), eyes sitting exactly on mesh/grid planes, eyes
    inside the scene box, and a batch of random cameras: the NaN/inf behaviour of
    every comparison must match the oracle's."""
    rng = np.random.default_rng(11)
    views = [
        dict(position=(0.0, 1.5, 0.0), forward=(0.0, -1.0, 1e-9)),     # straight down (forward.x == 0 exactly)
        dict(position=(0.0, 0.05, -1.0), forward=(0.0, 0.0, 1.0)),      # along +z in the x = 0 plane, on the box face
        dict(position=(-1.0, 0.1, 0.0), forward=(1.0, 0.0, 0.0)),       # along +x from the box face
        dict(position=(0.25, 0.0, 0.25), forward=(0.6, 0.0, 0.8)),      # horizontal, inside the box
        dict(position=(0.0, -0.5, 0.0), forward=(0.0, 1.0, 1e-9)),      # from below: every face is a back face
    ]
    for _ in range(6):
        f = rng.normal(size=3)
        views.append(dict(position=tuple(rng.uniform(-1.5, 1.5, 3)), forward=tuple(f / np.linalg.norm(f))))
    for name in ("hf22", "soup500"):
        scene, _ = scene_cache(name, sah=sah)
        renderer.set_meshes(scene)
        renderer.set_params(mode=1, depth=4)
        renderer.create_image(96, 64, aov=True)
        for k, v in enumerate(views):
            cam = clpt.cam_matrix(clpt.make_camera(near=0.1, far=1.0, fov=1.0, **v), 64)
            renderer.set_camera_matrix(cam)
            renderer.execute()
            img = renderer.read_image()
            prim, t, uv = renderer.read_aov()
            ref = oracle.render(scene, cam, 96, 64, mode=1, depth=4)
            assert np.array_equal(prim, ref["prim"]), (name, k)
            _assert_bit_equal(t, ref["t"], f"{name} view {k} t")
            _assert_bit_equal(img, ref["rgba"], f"{name} view {k} rgba")


def test_zero_matrix_and_far_camera(clpt, renderer, scene_cache):
    """Frame 0 of the reference runs with an unset matrix; a camera that misses the
    scene is all white (src/kernel.cl:421)."""
    scene, _ = scene_cache("hf22")
    renderer.set_meshes(scene)
    renderer.set_params(mode=1, depth=2)
    renderer.create_image(64, 48)
    renderer.set_camera_matrix(np.zeros((4, 4), dtype=np.float32))
    renderer.execute()
    assert np.isfinite(renderer.read_image()[..., 3]).all()
    cam = clpt.cam_matrix(clpt.make_camera(position=(0, 50, 0), forward=(0, 1, 0)), 48)
    renderer.set_camera_matrix(cam)
    renderer.execute()
    assert np.array_equal(renderer.read_image(), np.ones((48, 64, 4), dtype=np.float32))


def test_full_size_properties_1m(clpt, oracle, renderer):
    """BASELINE's full size (1080p, ~1M triangles, deep tree): oracle parity on a
    band of rows, plus size-independent properties over the whole frame."""
    from clpathtracer_b200 import scenes

    v, c, n = scenes.heightfield(707, False)
    # the tree bench.py renders: exact sweep over every triangle bound (--sah-bins 0), perfect splits
    scene = clpt.build_kd_sah(v, c, n, nbins=0, intersect_cost=1.0, empty_bonus=0.9)
    assert scene.stats()["nodes"] > 1_500_000  # (the binned builder makes about half as many)
    w, h = 1920, 1080
    cam = _cam(clpt, "canonical", h)
    # the bench parameters themselves (depth 5, jitter, several samples) on two bands: one under the
    # horizon (grazing rays, the costliest rows) and one near the bottom of the frame
    for band, spp in (((424, 440), 4), ((1000, 1008), 8)):
        kw = dict(mode=1, depth=5, spp=spp, seed=0, flags=clpt.FLAG_JITTER)
        img = _render_gpu(renderer, scene, cam, w, h, aov=False, **kw)[0]
        ref = oracle.render(scene, cam, w, h, rows=band, aov=False, **kw)
        _assert_bit_equal(img[slice(*band)], ref["rgba"][slice(*band)], f"bench tree, depth 5, {spp} spp, rows {band}")
    img, prim, t, uv = _render_gpu(renderer, scene, cam, w, h, mode=1, depth=2)
    band = (520, 552)
    ref = oracle.render(scene, cam, w, h, mode=1, depth=2, rows=band)
    sl = slice(*band)
    assert np.array_equal(prim[sl], ref["prim"][sl])
    _assert_bit_equal(t[sl], ref["t"][sl], "t")
    _assert_bit_equal(img[sl], ref["rgba"][sl], "rgba")
    # idempotence
    img2 = _render_gpu(renderer, scene, cam, w, h, mode=1, depth=2)[0]
    assert np.array_equal(img, img2)
    # every reported hit is a real intersection of that triangle at that t (recomputed in float64)
    ys, xs = np.nonzero(prim >= 0)
    pick = np.random.default_rng(0).choice(len(ys), 5000, replace=False)
    ys, xs = ys[pick], xs[pick]
    tri = scene.tris[:, 0].reshape(-1, 3)[prim[ys, xs]]
    p0, p1, p2 = (scene.verts[tri[:, k], :3].astype(np.float64) for k in range(3))
    bary = uv[ys, xs].astype(np.float64)
    hit = p0 + bary[:, :1] * (p1 - p0) + bary[:, 1:] * (p2 - p0)
    eye = (cam[:3, 2] / cam[3, 2]).astype(np.float64)
    dist = np.linalg.norm(hit - eye, axis=1)
    assert np.allclose(dist, t[ys, xs], rtol=1e-4, atol=1e-5)
    # depth-1 mirror of a depth-2 frame: mode A and mode B agree on which pixels hit
    a_img, a_prim, _, _ = _render_gpu(renderer, scene, cam, w, h, mode=0, depth=2)
    assert np.array_equal(a_prim, prim)
    assert np.array_equal(a_img[prim < 0], np.ones_like(a_img[prim < 0]))


def test_moving_camera_keeps_parity_while_the_claim_direction_adapts(clpt, oracle, renderer, scene_cache):
    """The claim direction (decided from the PREVIOUS frame's per-row cost, csrc/host/frame_sched.c)
    is a scheduling decision only: with a camera that tilts from looking down to looking at the
    horizon and back -- the costly band moves across the frame and flips sides -- every frame
    still equals the oracle's."""
    scene, _ = scene_cache("hf224", sah=True)
    w, h = 640, 360
    kw = dict(mode=1, depth=3, spp=2, seed=2, flags=clpt.FLAG_JITTER)
    renderer.set_meshes(scene)
    renderer.set_params(**kw)
    renderer.create_image(w, h)
    tilts = [-0.9, -0.9, -0.45, -0.45, -0.1, -0.1, 0.05, -0.6, -0.6]
    for i, ty in enumerate(tilts):
        fwd = np.array([0.0, ty, 0.9075])
        cam = clpt.cam_matrix(clpt.make_camera(near=0.1, far=1.0, fov=np.pi / 3, position=(0.0, 0.9, -1.7),
                                               forward=tuple(fwd / np.linalg.norm(fwd))), h)
        renderer.set_camera_matrix(cam)
        renderer.execute()
        ref = oracle.render(scene, cam, w, h, aov=False, **kw)["rgba"]
        _assert_bit_equal(renderer.read_image(), ref, f"frame {i} (tilt {ty})")


def _unorm8(frame):
    """What write_imagef stores into a CL_UNORM_INT8 image: clamp, x255 in fp32, round to nearest even."""
    return np.rint(np.clip(frame, 0.0, 1.0).astype(np.float32) * np.float32(255.0)).astype(np.uint8)


def test_readback_formats_and_pipeline(clpt, oracle, renderer, scene_cache):
    """CLReadImageRGBA8 = the float4 frame quantised like the reference's RGBA8 render target
    (src/GLHandler.c:177-185); CLReadImageAsync delivers the frame that was current when it was
    called although later frames have been rendered since; progressive frames are normalised."""
    scene, _ = scene_cache("hf22n")
    w, h = 333, 197
    cam = _cam(clpt, "canonical", h)
    frames = {}
    for seed in (1, 2, 3):
        frames[seed] = oracle.render(scene, cam, w, h, mode=1, depth=3, spp=2, seed=seed, flags=clpt.FLAG_JITTER,
                                     aov=False)["rgba"]
    img = _render_gpu(renderer, scene, cam, w, h, aov=False, mode=1, depth=3, spp=2, seed=1, flags=clpt.FLAG_JITTER)[0]
    _assert_bit_equal(img, frames[1], "float4")
    assert np.array_equal(renderer.read_image_rgba8(), _unorm8(frames[1]))
    # three reads in flight order: float4 of frame 1, RGBA8 of frame 2, RGBA8 of frame 3
    a = np.zeros((h, w, 4), dtype=np.float32)
    b = np.zeros((h, w, 4), dtype=np.uint8)
    c = np.zeros((h, w, 4), dtype=np.uint8)
    renderer.read_image_async(a)
    renderer.set_params(mode=1, depth=3, spp=2, seed=2, flags=clpt.FLAG_JITTER)
    renderer.execute()
    renderer.read_image_async(b)
    renderer.read_wait(1)            # the first read has landed, the second may still be in flight
    _assert_bit_equal(a, frames[1], "async float4")
    renderer.set_params(mode=1, depth=3, spp=2, seed=3, flags=clpt.FLAG_JITTER)
    renderer.execute()
    renderer.read_image_async(c)     # reuses the first staging buffer
    renderer.read_wait(0)
    assert np.array_equal(b, _unorm8(frames[2])) and np.array_equal(c, _unorm8(frames[3]))
    # progressive: the RGBA8 read-back is the normalised running mean
    renderer.set_params(mode=1, depth=3, spp=1, seed=4, flags=clpt.FLAG_JITTER | clpt.FLAG_ACCUMULATE)
    renderer.create_image(w, h)
    acc = oracle.new_accumulator(w, h)
    for base in range(3):
        renderer.execute()
        mean = oracle.render(scene, cam, w, h, mode=1, depth=3, spp=1, seed=4, aov=False, sample_base=base,
                             flags=oracle.FLAG_JITTER | oracle.FLAG_ACCUMULATE, accumulate_into=acc)["rgba"]
    _assert_bit_equal(renderer.read_image(), mean, "progressive float4")
    assert np.array_equal(renderer.read_image_rgba8(), _unorm8(mean))
    renderer.set_params(mode=0, depth=2)


def test_progressive_sharded_by_sample(clpt, oracle, renderer, scene_cache):
    """Progressive accumulation across ranks is spread by SAMPLE: rank r
    of N renders the WHOLE frame with samples base + r*spp .. of every round of N*spp samples into
    its own fixed-point sums; the read-back adds the ranks' sums (here, without a communicator,
    each rank's own).  The three ranks' sums added together are the single-rank sums of the same
    eighteen samples, bit for bit."""
    scene, _ = scene_cache("hf22n")
    w, h = 200, 150
    cam = _cam(clpt, "canonical", h)
    L = clpt.lib()
    kw = dict(mode=1, depth=3, spp=2, seed=6)
    oflags = oracle.FLAG_JITTER | oracle.FLAG_ACCUMULATE
    try:
        for rank in range(3):
            renderer.set_meshes(scene)
            renderer.set_camera_matrix(cam)
            renderer.set_params(flags=clpt.FLAG_JITTER | clpt.FLAG_ACCUMULATE, **kw)
            renderer.create_image(w, h)
            L.CLSetTileShard(rank, 3, 8)
            acc = oracle.new_accumulator(w, h)
            for frame in range(3):
                renderer.execute()
                mean = oracle.render(scene, cam, w, h, aov=False, sample_base=frame * 6 + rank * 2, accumulate_into=acc,
                                     flags=oflags, **kw)["rgba"]
            _assert_bit_equal(renderer.read_image(), mean, f"rank {rank}")
            assert np.array_equal(renderer.read_image_rgba8(), _unorm8(mean))
    finally:
        L.CLSetTileShard(0, 1, 8)
        renderer.set_params(mode=0, depth=2)


def test_band_parity_10m(clpt, oracle, renderer):
    """BASELINE config 4's scene: 9,999,392 triangles, SAH tree (exact sweep), 3840x2160, one
    jittered sample, depth 5 -- bit parity with the oracle on a band of rows, and the
    progressive accumulation of two frames against the oracle's."""
    from clpathtracer_b200 import scenes

    v, c, n = scenes.heightfield(2236, False)
    scene = clpt.build_kd_sah(v, c, n, nbins=0, intersect_cost=1.0, empty_bonus=0.9)
    assert scene.n_tris == 9_999_392
    w, h = 3840, 2160
    cam = _cam(clpt, "canonical", h)
    band = (848, 864)
    sl = slice(*band)
    kw = dict(mode=1, depth=5, spp=1, seed=0)
    img = _render_gpu(renderer, scene, cam, w, h, aov=False, flags=clpt.FLAG_JITTER, **kw)[0]
    ref = oracle.render(scene, cam, w, h, rows=band, aov=False, flags=oracle.FLAG_JITTER, **kw)
    _assert_bit_equal(img[sl], ref["rgba"][sl], "10M triangles, 4K, depth 5")
    assert (img[..., :3] != 1.0).any(axis=-1).mean() > 0.2  # the camera sees the terrain
    # progressive: two 1-spp frames accumulate to the oracle's two-frame sums; the read-back is the mean
    renderer.set_params(flags=clpt.FLAG_JITTER | clpt.FLAG_ACCUMULATE, **kw)
    renderer.create_image(w, h)
    renderer.execute()
    renderer.execute()
    got = renderer.read_image()
    acc = oracle.new_accumulator(w, h)
    for base in (0, 1):
        want = oracle.render(scene, cam, w, h, rows=band, aov=False, flags=oracle.FLAG_JITTER | oracle.FLAG_ACCUMULATE,
                             sample_base=base, accumulate_into=acc, **kw)["rgba"]
    _assert_bit_equal(got[sl], want[sl], "two accumulated frames")
    renderer.set_params(mode=0, depth=2)


def test_fails_loudly(clpt):
    """Misuse aborts with a message, like the reference's HANDLE_ERR (src/error.c:147-154)."""
    root = Path(__file__).resolve().parents[1]
    code = ("import sys;sys.path.insert(0,%r);import clpathtracer_b200 as cl;L=cl.lib();%s")
    for body, needle in [("L.CLExecute(64,64)", "CLInit has not been called"),
                         ("L.CLInit(None,None);L.CLExecute(64,64)", "no render target"),
                         ("L.CLInit(None,None);L.CLCreateImageHeadless(8,8);L.CLExecute(8,8)", "no scene"),
                         ("L.CLInit(None,None);L.CLCreateImage(3)", "CUDA Error"),  # no GL context on the box
                         ("L.CLInit(b'k.cl',b'trace')", "no kernel named"),
                         # the round-2 additions abort the same way
                         ("L.CLInit(None,None);L.CLCreateImageHeadless(8,8);L.CLReadImageRGBA8(None,7)", "bytes given"),
                         ("L.CLInit(None,None);L.CLCreateImageHeadless(8,8);L.CLReadImageAsync(None,256,5)", "format must be"),
                         ("L.CLInit(None,None);L.CLUpdateVertices(0,None,16)", "no mesh was uploaded"),
                         ("L.CLInit(None,None);L.CLRebuildMeshes()", "no mesh was uploaded"),
                         ("L.CLInit(None,None);k=cl.KD();L.CLDownloadKd(k)", "not built by CLBuildMeshes"),
                         ("L.CLInit(None,None);L.CLSetEngine(3)", "CLSetEngine")]:
        p = subprocess.run([sys.executable, "-c", code % (str(root), body)], capture_output=True, text=True)
        assert p.returncode == 1, (body, p.returncode, p.stderr)
        assert needle in p.stderr, (body, p.stderr)
    # inconsistent scenes are rejected on upload, before any kernel can dereference them
    prep = ("from clpathtracer_b200 import scenes;s=cl.build_kd(*scenes.heightfield(6,True));"
            "r=cl.Renderer(0);")
    for corrupt, needle in [("s.nodes['c'][0,1]=10**6", "malformed"), ("s.tri_indices[3]=10**6", "tri_indices"),
                            ("s.tris[5,0]=10**6", "missing vertex"), ("s.tris[0,1]=10**6", "normal"),
                            ("s.nodes['type'][2]=7", "has type"),
                            ("i=int((s.nodes['type']==1).argmax());s.nodes['c'][i,2]=10**6", "rope")]:
        p = subprocess.run([sys.executable, "-c", code % (str(root), prep + corrupt + ";r.set_meshes(s)")],
                           capture_output=True, text=True)
        assert p.returncode == 1, (corrupt, p.returncode, p.stderr)
        assert "invalid scene" in p.stderr and needle in p.stderr, (corrupt, p.stderr)
