import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def clpt():
    """The package with libclpt.so built (nvcc cross-compiles without a GPU)."""
    import clpathtracer_b200 as cl
    from clpathtracer_b200 import build as b

    b.build()
    cl.lib()
    return cl


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as op

    op.build()
    op.oracle()
    return op


@pytest.fixture(scope="session")
def scene_cache(clpt):
    """Scenes built once per session: name -> (Scene, extras)."""
    from clpathtracer_b200 import scenes

    cache = {}

    def get(name, depth=15, nbins=25, sah=False):
        key = (name, depth, nbins, sah)
        if key in cache:
            return cache[key]
        extras = {}
        if name.startswith("hf"):
            # hf22n = heightfield n=22 with normals, hf22 = without
            with_n = name.endswith("n")
            n = int(name[2:-1] if with_n else name[2:])
            v, c, nn = scenes.heightfield(n, with_n)
        elif name == "cornell":
            v, c, nn, tm = scenes.cornell(10)
            extras["tri_material"] = tm
        elif name.startswith("soup"):
            v, c, nn = scenes.soup(int(name[4:]))
        else:
            raise KeyError(name)
        if sah:
            # sah="exact": candidate planes on every triangle bound at every cell size
            s = clpt.build_kd_sah(v, c, nn, nbins=0 if sah == "exact" else 32, intersect_cost=1.0, empty_bonus=0.9,
                                  clip=sah != "noclip")
        else:
            s = clpt.build_kd(v, c, nn, depth=depth, nbins=nbins)
        cache[key] = (s, extras)
        return cache[key]

    return get


@pytest.fixture(scope="session")
def renderer(clpt):
    """The process-wide renderer (the library state is a singleton)."""
    import torch  # noqa: F401  (only to fail early and clearly without CUDA)

    r = clpt.Renderer(device=int(os.environ.get("CLPT_DEVICE", "0")))
    yield r
    r.close()
