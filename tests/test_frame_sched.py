"""Host-side scheduling rules of the render path (csrc/host/frame_sched.c), driven
through the library without a GPU: the claim direction of the persistent render
kernel, decided from the previous frame's per-row cost."""
import ctypes as C

import numpy as np


def _direction(clpt, cost, current):
    L = clpt.lib()
    L.clpt_claim_direction.restype = C.c_int
    L.clpt_claim_direction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    cost = np.ascontiguousarray(cost, dtype=np.uint64)
    where = C.c_double(-1.0)
    d = L.clpt_claim_direction(cost.ctypes.data, len(cost), current, C.byref(where))
    return d, where.value


def _frame(rows, horizon, sky_above=True):
    """Per-row cost of a ground/sky frame: cheap sky, a costly band of grazing rows
    right under the horizon, moderate near-field ground."""
    cost = np.full(rows, 1000, dtype=np.uint64)          # near field
    band = slice(max(horizon - 12, 0), horizon)
    cost[band] = 20000                                   # grazing rows
    cost[horizon:] = 3                                   # sky
    return cost if sky_above else cost[::-1].copy()


def test_costly_band_in_the_far_half_flips_the_direction(clpt):
    cost = _frame(540, 320)                              # band at ~0.58 of the way, sky after it
    d, where = _direction(clpt, cost, 0)
    assert d == 1 and 0.55 < where < 0.62                # start from the far end: the band comes early
    assert _direction(clpt, cost, 1)[0] == 1             # and stays


def test_mirrored_frame_keeps_or_restores_top_down(clpt):
    cost = _frame(540, 320, sky_above=False)             # band at ~0.42
    assert _direction(clpt, cost, 0)[0] == 0
    assert _direction(clpt, cost, 1)[0] == 0


def test_hysteresis_and_degenerate_inputs(clpt):
    rows = 200
    cost = np.full(rows, 10, dtype=np.uint64)
    cost[98:103] = 1000                                  # band at the centre: inside the dead zone
    assert _direction(clpt, cost, 0)[0] == 0 and _direction(clpt, cost, 1)[0] == 1
    assert _direction(clpt, np.zeros(rows, dtype=np.uint64), 1)[0] == 1     # nothing measured
    assert _direction(clpt, np.arange(8, dtype=np.uint64), 0)[0] == 0       # too few rows to judge
    rng = np.random.default_rng(5)
    flat = rng.integers(900, 1100, size=rows).astype(np.uint64)             # noise only: wherever the
    d0 = _direction(clpt, flat, 0)[0]                                       # maximum falls, the answer
    assert _direction(clpt, flat, d0)[0] == d0                              # is stable once taken

