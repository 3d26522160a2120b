"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

DECL = re.compile(r"^[A-Za-z_][\w \*]*?[\s\*]([A-Za-z_]\w*)\s*\(", re.M)
NOT_FUNCTIONS = {"defined", "sizeof", "CLPT_STATIC_ASSERT", "__attribute__", "Vector3", "Vector4",
                 "vector_append", "vector_length", "HANDLE_ERR", "vec_x", "vec_y", "vec_z"}


def declared_functions(header: Path) -> set[str]:
    text = header.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # comments
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"^\s*#[^\n]*(\\\n[^\n]*)*", "", text, flags=re.M)  # preprocessor lines (with continuations)
    text = text.replace('extern "C" {', "")
    names = set()
    for stmt in text.split(";"):
        stmt = stmt.strip()
        if "(" not in stmt or stmt.startswith("typedef") or "{" in stmt:
            continue
        m = DECL.search(stmt)
        if m and m.group(1) not in NOT_FUNCTIONS:
            names.add(m.group(1))
    return names


def test_headers_declare_the_reference_boundary():
    """The eight CLState.h entry points of the reference, by name (include/CLState.h:13-28)."""
    names = declared_functions(ROOT / "include" / "CLState.h")
    for fn in ["CLInit", "CLTerminate", "CLSetCameraMatrix", "CLSetObjects", "CLSetMeshes", "CLDeleteImage",
               "CLCreateImage", "CLExecute"]:
        assert fn in names
    handler = declared_functions(ROOT / "include" / "CLHandler.h")
    for fn in ["CLGetPlatform", "CLGetDevice", "CLCreateContext", "CLBuildProgram", "CLCreateQueue",
               "CLCreateKernel", "CLCreateBuffer", "CLEnqueueKernel", "err_string", "handle_err"]:
        assert fn in handler
    host = declared_functions(ROOT / "include" / "clpt_host.h")
    for fn in ["new_list", "init_list", "copy_list", "delete_list", "list_grow", "list_size", "list_concat",
               "vec_dot", "vec_cross", "vec_normalize", "mat_multiply", "mat_inverse", "mat_set", "mat_get",
               "cam_matrix", "build_kd", "parse_kd", "delete_kd", "LoadModel", "AddPhysObject", "PhysStep"]:
        assert fn in host


def test_library_exports_every_declared_symbol(clpt):
    lib = C.CDLL(str(clpt.LIB_PATH))
    missing = []
    total = 0
    for header in sorted((ROOT / "include").glob("*.h")):
        for name in sorted(declared_functions(header)):
            total += 1
            try:
                getattr(lib, name)
            except AttributeError:
                missing.append(f"{header.name}:{name}")
    assert total > 60
    assert not missing, missing


def test_no_cpu_fallback_in_product():
    """Nothing under the package loads, links, imports or includes anything from
    oracle/ (comments may mention it), and the Python side has no alternative to
    the CUDA library."""
    forbidden = [re.compile(p) for p in (r"liboracle", r"oracle_py", r"from\s+oracle", r"import\s+oracle",
                                         r"#include\s+[\"<][^\n]*oracle", r"libref_", r"_ref/")]
    for path in (ROOT / "clpathtracer_b200").rglob("*"):
        if path.suffix in {".py", ".c", ".cu", ".cpp", ".h", ".cuh"} and "_build" not in path.parts \
                and path.name != "build.py":  # build.py runs `make -C oracle`: builds the checker, never loads it
            text = path.read_text()
            for pat in forbidden:
                assert not pat.search(text), (path, pat.pattern)
    init = (ROOT / "clpathtracer_b200" / "__init__.py").read_text()
    assert "There is no CPU or PyTorch fallback" in init
    build_py = (ROOT / "clpathtracer_b200" / "build.py").read_text()
    assert "oracle" in build_py  # build() may BUILD the checker; it never loads it
