"""Host half of the path: lists, vec/matrix/camera, kd builder, .kd / OBJ I/O.

Pinned two ways:
  * against tests/golden/host_golden.json, produced by the reference's own
    unmodified host code (tests/golden/make_golden.py) -- runs everywhere;
  * live against oracle/_ref/libref_host.so when it exists (the build
    container, or a box the prebuilt .so travelled to).
Bar: byte-exact (SURVEY.md section 8d parity metric (i)).
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

GOLDEN = json.loads((Path(__file__).parent / "golden" / "host_golden.json").read_text())


def _scene_inputs(name):
    from clpathtracer_b200 import scenes

    if name.startswith("hf"):
        with_n = name.endswith("n")
        n = int(name[2:-1] if with_n else name[2:])
        return scenes.heightfield(n, with_n)
    if name == "cornell":
        return scenes.cornell(10)[:3]
    if name.startswith("soup"):
        return scenes.soup(int(name[4:]))
    raise KeyError(name)


# ------------------------------------------------------------------ wire types
def test_wire_sizes(clpt):
    assert clpt.KDNODE_DTYPE.itemsize == 68
    assert C.sizeof(clpt.Matrix) == 64
    assert C.sizeof(clpt.Camera) == 48
    assert C.sizeof(clpt.KD) == 40
    assert C.sizeof(clpt.CLMaterial) == 32
    f = clpt.KDNODE_DTYPE.fields
    assert (f["type"][1], f["a"][1], f["b"][1], f["c"][1]) == (32, 36, 40, 44)


def test_list_semantics(clpt):
    """include/list.h: byte lengths, list_grow index, NULL delete, copy."""
    L = clpt.lib()
    p = L.new_list(0)
    assert L.list_size(p) == 0
    L.list_grow.restype, L.list_grow.argtypes = C.c_size_t, [C.POINTER(C.c_void_p), C.c_size_t]
    ref = C.c_void_p(p)
    for k in range(100):
        assert L.list_grow(C.byref(ref), 12) == k  # old_length / size
        C.memmove(ref.value + 12 * k, bytes([k]) * 12, 12)
    assert L.list_size(ref) == 1200
    q = L.copy_list(ref)
    assert C.string_at(q, 1200) == C.string_at(ref.value, 1200)
    r = L.init_list(7, 16)
    assert L.list_size(r) == 112
    for x in (ref, q, r):
        L.delete_list(x)
    L.delete_list(None)
    a = np.arange(10, dtype=np.int32)
    assert np.array_equal(clpt.from_list(clpt.to_list(a), np.int32), a)


# ------------------------------------------------------------------ camera
@pytest.mark.parametrize("name", sorted(GOLDEN["cam"]))
def test_cam_matrix_golden(clpt, name):
    g = GOLDEN["cam"][name]
    kw = dict(g["camera"])
    m = clpt.cam_matrix(clpt.make_camera(**kw), g["height"])
    assert m.view(np.uint32).reshape(-1).tolist() == g["matrix_bits"]


def test_cam_matrix_known_answer(clpt):
    """the reference's start-up camera at height 480."""
    from clpathtracer_b200 import scenes

    m = clpt.cam_matrix(clpt.make_camera(**scenes.REFERENCE_CAMERA), 480)
    want = np.array([[0.002405626, 0, 0, 0], [0, 0.002405626, -0.449999988, 0.550000012],
                     [0, 0, 0.899999917, -0.100000054], [0, 0, -4.49999952, 5.5]], dtype=np.float32)
    assert np.array_equal(m, want)
    eye = m[:3, 2] / m[3, 2]  # src/kernel.cl:443-445
    assert np.allclose(eye, [0, 0.1, -0.2], atol=1e-6)


def test_cam_matrix_vs_reference_random(clpt, oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    R = oracle.ref()
    rng = np.random.default_rng(7)
    for _ in range(200):
        fwd = rng.normal(size=3)
        fwd /= np.linalg.norm(fwd)
        cam = clpt.make_camera(near=rng.uniform(0.01, 1), far=rng.uniform(1.5, 50), fov=rng.uniform(0.3, 2.5),
                               position=rng.uniform(-3, 3, 3), forward=fwd)
        h = int(rng.integers(16, 2200))
        mine = clpt.cam_matrix(cam, h)
        m = clpt.Matrix()
        R.ref_cam_matrix_ptr(C.byref(cam), h, C.byref(m))
        theirs = np.frombuffer(bytes(m), dtype=np.float32).reshape(4, 4)
        assert np.array_equal(mine.view(np.uint32), theirs.view(np.uint32))


def test_matrix_ops_vs_reference(clpt, oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    R, L = oracle.ref(), clpt.lib()
    M = clpt.Matrix
    for lib_ in (R, L):
        lib_.mat_multiply.restype, lib_.mat_multiply.argtypes = M, [M, M]
        lib_.mat_inverse.restype, lib_.mat_inverse.argtypes = M, [M, C.POINTER(C.c_int)]
        lib_.mat_add.restype, lib_.mat_add.argtypes = M, [M, M]
    rng = np.random.default_rng(3)

    def mk(a):
        m = M()
        C.memmove(C.byref(m), np.ascontiguousarray(a, dtype=np.float32).ctypes.data, 64)
        return m

    for k in range(100):
        a, b = rng.normal(size=(4, 4)), rng.normal(size=(4, 4))
        if k == 0:
            a = np.zeros((4, 4))  # singular -> zero matrix and err = 1
        for fn in ("mat_multiply", "mat_add"):
            assert bytes(getattr(L, fn)(mk(a), mk(b))) == bytes(getattr(R, fn)(mk(a), mk(b)))
        e1, e2 = C.c_int(0), C.c_int(0)
        assert bytes(L.mat_inverse(mk(a), C.byref(e1))) == bytes(R.mat_inverse(mk(a), C.byref(e2)))
        assert e1.value == e2.value == (1 if k == 0 else 0)


# ------------------------------------------------------------------ kd builder
@pytest.mark.parametrize("name", sorted(GOLDEN["kd"]))
def test_build_kd_golden(clpt, name):
    """build_kd at depth 15 / 25 bins is byte-identical to the reference builder."""
    g = GOLDEN["kd"][name]
    v, c, n = _scene_inputs(name)
    assert hashlib.sha256(clpt._as_vec4(v).tobytes()).hexdigest() == g["verts_sha256"], "scene generator drifted"
    s = clpt.build_kd(v, c, n)
    assert len(s.nodes) == g["nodes"] and len(s.tri_indices) == g["tri_refs"]
    assert hashlib.sha256(s.nodes.tobytes()).hexdigest() == g["nodes_sha256"]
    assert hashlib.sha256(s.tri_indices.tobytes()).hexdigest() == g["tri_indices_sha256"]


def test_survey_known_answers(clpt):
    """SURVEY.md section 4 KAT table (reference builder, n = 22 heightfield)."""
    from clpathtracer_b200 import scenes

    s = clpt.build_kd(*scenes.heightfield(22, True))
    st = s.stats()
    assert (st["nodes"], st["leaves"], st["empty_leaves"], st["leaf_tri_refs"], st["max_leaf_tris"]) == \
        (7321, 3661, 1488, 7756, 8)


def _ref_build(oracle, clpt, v, c, n):
    R = oracle.ref()

    def mk(a):
        if a is None:
            return R.new_list(0)
        a = np.ascontiguousarray(a)
        p = R.init_list(a.nbytes, 1)
        C.memmove(p, a.ctypes.data, a.nbytes)
        return p

    R.build_kd.restype, R.build_kd.argtypes = clpt.KD, [C.c_void_p] * 3 + [C.c_char_p]
    k = R.build_kd(mk(c), mk(clpt._as_vec4(v)), mk(clpt._as_vec4(n) if n is not None else None), None)
    return C.string_at(k.node_vec, R.list_size(k.node_vec)), C.string_at(k.tri_indices, R.list_size(k.tri_indices))


@pytest.mark.parametrize("name", ["hf4n", "hf22", "cornell", "soup500", "hf33", "soup3000"])
def test_build_kd_vs_reference_live(clpt, oracle, name):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    v, c, n = _scene_inputs(name)
    s = clpt.build_kd(v, c, n)
    nodes, idx = _ref_build(oracle, clpt, v, c, n)
    assert s.nodes.tobytes() == nodes
    assert s.tri_indices.tobytes() == idx


def test_build_kd_degenerate_inputs(clpt):
    """Edge cases: one triangle, a flat (zero-extent) scene, duplicate triangles, no triangles."""
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1]], dtype=np.float32)
    one = clpt.build_kd(v, clpt.corners_from_faces(np.array([[0, 2, 1]]), False))
    assert len(one.nodes) == 1 and one.nodes["type"][0] == 1 and one.nodes["b"][0] == 1
    assert one.nodes["c"][0].tolist() == [-1] * 6
    dup = clpt.build_kd(v, clpt.corners_from_faces(np.array([[0, 2, 1]] * 5), False))
    assert dup.stats()["leaf_tri_refs"] >= 5
    none = clpt.build_kd(v, np.zeros((0, 4), dtype=np.int32))
    assert len(none.nodes) == 1 and none.nodes["b"][0] == 0


def _check_tree_invariants(s):
    nodes = s.nodes
    n = len(nodes)
    split = nodes["type"] == 0
    # preorder: the left child is always parent + 1 (src/kd_tree.c:187-199)
    idx = np.arange(n)
    assert np.array_equal(nodes["c"][split, 0], idx[split] + 1)
    assert np.all(nodes["c"][split, 1] > nodes["c"][split, 0])
    leaf = ~split
    first, cnt = nodes["a"][leaf], nodes["b"][leaf]
    # leaves tile tri_indices contiguously in preorder
    assert np.array_equal(first, np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    assert cnt.sum() == len(s.tri_indices)
    # every triangle is referenced somewhere
    assert np.array_equal(np.unique(s.tri_indices), np.arange(s.n_tris))
    # children partition the parent's box at the plane
    for i in np.flatnonzero(split)[:2000]:
        ax = nodes["b"][i]
        plane = nodes["a"][i:i + 1].view(np.float32)[0]
        l, r = nodes["c"][i, 0], nodes["c"][i, 1]
        assert nodes["max"][l, ax] == plane and nodes["min"][r, ax] == plane
        assert nodes["min"][i, ax] < plane < nodes["max"][i, ax]
    # ropes: -1 or a node whose box touches the face and covers the leaf's extent there
    ropes = nodes["c"][leaf]
    assert np.all((ropes >= -1) & (ropes < n))
    li = np.flatnonzero(leaf)
    for i in li[:: max(1, len(li) // 500)]:
        for f in range(6):
            r = nodes["c"][i, f]
            ax, hi = f // 2, f % 2
            if r < 0:
                continue
            face = nodes["max"][i, ax] if hi else nodes["min"][i, ax]
            other = nodes["min"][r, ax] if hi else nodes["max"][r, ax]
            assert other == face
            for a2 in range(3):
                if a2 != ax:
                    assert nodes["min"][r, a2] <= nodes["min"][i, a2] and nodes["max"][r, a2] >= nodes["max"][i, a2]


@pytest.mark.parametrize("depth,nbins", [(15, 25), (8, 7), (20, 25), (22, 31)])
def test_build_kd_ex_invariants(clpt, depth, nbins):
    from clpathtracer_b200 import scenes

    v, c, n = scenes.heightfield(40, False)
    s = clpt.build_kd(v, c, n, depth=depth, nbins=nbins)
    _check_tree_invariants(s)
    st = s.stats()
    assert st["leaves"] * 2 - 1 == st["nodes"]


@pytest.mark.parametrize("variant", [dict(), dict(nbins=0), dict(clip=False), dict(nbins=0, clip=False)],
                         ids=["binned", "exact", "binned-noclip", "exact-noclip"])
@pytest.mark.parametrize("name", ["hf40", "cornell", "soup3000", "hf4n"])
def test_build_kd_sah_invariants(clpt, name, variant):
    """The SAH builder (extension) emits the same wire format: preorder nodes,
    contiguous leaf runs, every triangle referenced, valid ropes -- with binned or
    exact-sweep candidate planes, with and without perfect splits."""
    v, c, n = _scene_inputs(name)
    s = clpt.build_kd_sah(v, c, n, intersect_cost=1.0, empty_bonus=0.9, **variant)
    _check_tree_invariants(s)
    st = s.stats()
    assert st["leaves"] * 2 - 1 == st["nodes"]
    ref = clpt.build_kd(v, c, n)
    # it terminates by cost, not only by depth: far fewer references than the reference heuristic on meshes
    if name == "hf40":
        assert st["leaf_tri_refs"] < ref.stats()["leaf_tri_refs"]
        assert st["max_leaf_tris"] <= 16


def test_build_kd_thread_independent(clpt):
    """The parallel evaluation order must not change a single byte."""
    code = ("import sys,hashlib;sys.path.insert(0,%r);import clpathtracer_b200 as cl;"
            "from clpathtracer_b200 import scenes;s=cl.build_kd(*scenes.heightfield(150,False),depth=18);"
            "t=cl.build_kd_sah(*scenes.heightfield(150,False));u=cl.build_kd_sah(*scenes.soup(20000),nbins=0);"
            "print(hashlib.sha256(s.nodes.tobytes()+s.tri_indices.tobytes()+t.nodes.tobytes()+t.tri_indices.tobytes()"
            "+u.nodes.tobytes()+u.tri_indices.tobytes()).hexdigest())") % str(
                Path(__file__).resolve().parents[1])
    digests = set()
    for threads in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=threads)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True)
        digests.add(out.stdout.strip().splitlines()[-1])
    assert len(digests) == 1


# ------------------------------------------------------------------ .kd and OBJ
def test_kd_file_roundtrip(clpt, tmp_path):
    from clpathtracer_b200 import scenes

    v, c, n = scenes.heightfield(12, True)
    stem = str(tmp_path / "hf12")
    s = clpt.build_kd(v, c, n, path=stem)  # also writes <stem>.kd like the reference
    back = clpt.load_model(stem + ".kd")
    for a, b in [(s.nodes, back.nodes), (s.tri_indices, back.tri_indices), (s.tris, back.tris),
                 (s.verts, back.verts), (s.norms, back.norms)]:
        assert a.tobytes() == b.tobytes()
    # layout: [size_t n][n x 68]...  (src/kd_tree.c:250-271)
    raw = (tmp_path / "hf12.kd").read_bytes()
    assert int.from_bytes(raw[:8], "little") == len(s.nodes)
    assert raw[8:8 + 68 * len(s.nodes)] == s.nodes.tobytes()
    assert len(raw) == 5 * 8 + s.nodes.nbytes + s.verts.nbytes + s.norms.nbytes + s.tri_indices.nbytes + s.tris.nbytes
    # truncated file is an error, not garbage
    (tmp_path / "bad.kd").write_bytes(raw[: len(raw) // 2])
    with pytest.raises(RuntimeError):
        clpt.load_model(str(tmp_path / "bad.kd"))
    with pytest.raises(RuntimeError):
        clpt.load_model(str(tmp_path / "missing.kd"))
    with pytest.raises(RuntimeError):
        clpt.load_model(str(tmp_path / "model.stl"))


def test_kd_file_matches_reference_writer(clpt, oracle, tmp_path):
    """The .kd bytes written here equal the bytes the reference's build_kd writes,
    and the reference's parse_kd reads ours back."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    from clpathtracer_b200 import scenes

    v, c, n = scenes.heightfield(10, True)
    clpt.build_kd(v, c, n, path=str(tmp_path / "mine"))
    R = oracle.ref()

    def mk(a):
        a = np.ascontiguousarray(a)
        p = R.init_list(a.nbytes, 1)
        C.memmove(p, a.ctypes.data, a.nbytes)
        return p

    R.build_kd.restype, R.build_kd.argtypes = clpt.KD, [C.c_void_p] * 3 + [C.c_char_p]
    R.build_kd(mk(c), mk(clpt._as_vec4(v)), mk(clpt._as_vec4(n)), str(tmp_path / "ref").encode())
    assert (tmp_path / "mine.kd").read_bytes() == (tmp_path / "ref.kd").read_bytes()
    k = clpt.KD()
    R.parse_kd.restype, R.parse_kd.argtypes = C.c_int, [C.c_char_p, C.POINTER(clpt.KD)]
    assert R.parse_kd(str(tmp_path / "mine.kd").encode(), C.byref(k)) == 0
    assert R.list_size(k.node_vec) % 68 == 0


@pytest.mark.parametrize("with_normals", [True, False])
def test_obj_loader(clpt, oracle, tmp_path, with_normals):
    """OBJ text -> lists: equal to the generator's arrays, and (live) to what the
    reference's LoadModel builds from the same file."""
    from clpathtracer_b200 import scenes

    v, c, n = scenes.heightfield(9, with_normals)
    path = str(tmp_path / "m.obj")
    scenes.write_obj_text(path, v, c, n)
    s = clpt.load_model(path)
    assert np.array_equal(s.verts[:, :3], v) and np.all(s.verts[:, 3] == 0)
    assert np.array_equal(s.tris[:, 0], c[:, 0])
    assert np.all((s.tris[:, 1] >= 0) == with_normals)
    assert (tmp_path / "m.kd").exists()  # cached like src/model.c:134
    direct = clpt.build_kd(v, c, n)
    assert s.nodes.tobytes() == direct.nodes.tobytes()
    if oracle.have_ref():
        R = oracle.ref()
        k = clpt.KD()
        R.LoadModel.restype, R.LoadModel.argtypes = C.c_int, [C.c_char_p, C.POINTER(clpt.KD)]
        path2 = str(tmp_path / "r.obj")
        scenes.write_obj_text(path2, v, c, n)
        assert R.LoadModel(path2.encode(), C.byref(k)) == 0
        for mine, theirs in [(s.nodes, k.node_vec), (s.tri_indices, k.tri_indices), (s.verts, k.vert_vec),
                             (s.norms, k.norm_vec), (s.tris, k.tri_vec)]:
            assert mine.tobytes() == C.string_at(theirs, R.list_size(theirs))


def test_obj_polygons_and_relative_indices(clpt, tmp_path):
    p = tmp_path / "quad.obj"
    p.write_text("# quad as one polygon, relative indices, vt present\n"
                 "v 0 0 0\nv 1 0 0\nv 1 0 1\nv 0 0 1\nvt 0 0\nvn 0 1 0\n"
                 "f -4/1/1 -1/1/1 -2/1/1 -3/1/1\n")
    s = clpt.load_model(str(p))
    assert s.n_tris == 2  # fan: (0,3,2) (0,2,1)
    assert s.tris[:, 0].tolist() == [0, 3, 2, 0, 2, 1]
    assert s.tris[:, 1].tolist() == [0] * 6 and s.tris[:, 2].tolist() == [0] * 6


def test_physics_step(clpt):
    """src/physics.c:49-53: pos += vel * (float)dt."""
    L = clpt.lib()
    pos, vel = clpt.Vector4(), clpt.Vector4()
    pos.s[:] = [0.0, 0.1, -0.2, 0.0]
    vel.s[:] = [1.0, 0.0, 0.5, 0.0]
    L.AddPhysObject(C.byref(pos), C.byref(vel))
    for _ in range(3):
        L.PhysStep(0.016)
    dt = np.float32(0.016)
    want = np.array([0.0, 0.1, -0.2], dtype=np.float32)
    for _ in range(3):
        want = want + np.array([1.0, 0.0, 0.5], dtype=np.float32) * dt
    assert np.array_equal(np.array(pos.s[:3], dtype=np.float32), want)
    L.PhysTerminate()
