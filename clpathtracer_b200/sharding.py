"""Row-tile sharding of a frame across ranks.

Tiles of `tile_rows` image rows are dealt round-robin: tile t belongs to rank
t % nranks and is that rank's local tile t // nranks.  Each rank renders its
tiles into a compact slab of ceil(tiles / nranks) * tile_rows rows and, with
direct placement, straight into row slab_row -> image row of every rank's frame;
on the fallback path the slabs are all-gathered and de-interleaved.  This module
is the host-side statement of
that mapping; csrc/cuda/clstate.cu (slab_rows_for / local_rows_for) and the
kernels in render_kernel.cu implement the same arithmetic on the device.
"""
from __future__ import annotations

import numpy as np


def first_sample_of_rank(frame: int, rank: int, nranks: int, spp: int) -> int:
    """Progressive frames (CLPT_FLAG_ACCUMULATE) are spread over ranks by SAMPLE, not by region:
    frame k on N ranks adds samples [k*N*spp, (k+1)*N*spp) of every pixel, of which rank r
    renders [k*N*spp + r*spp, ... + spp) -- the whole frame -- into its own fixed-point sums;
    a read-back adds the ranks' sums (clstate.cu: clpt_state_launch_frame / displayable_frame)."""
    return (frame * nranks + rank) * spp


def slab_rows(height: int, nranks: int, tile_rows: int) -> int:
    tiles = -(-height // tile_rows)
    return -(-tiles // nranks) * tile_rows


def rows_of_rank(height: int, rank: int, nranks: int, tile_rows: int) -> np.ndarray:
    y = np.arange(height)
    return y[(y // tile_rows) % nranks == rank]


def slab_row_of(y: np.ndarray, nranks: int, tile_rows: int) -> np.ndarray:
    """Row index inside the owning rank's slab for image row(s) y."""
    t = y // tile_rows
    return (t // nranks) * tile_rows + (y - t * tile_rows)


def to_slab(image_rows: np.ndarray, height: int, rank: int, nranks: int, tile_rows: int) -> np.ndarray:
    """Pack a full-height image (only this rank's rows need be valid) into its slab."""
    rows = rows_of_rank(height, rank, nranks, tile_rows)
    slab = np.zeros((slab_rows(height, nranks, tile_rows),) + image_rows.shape[1:], dtype=image_rows.dtype)
    slab[slab_row_of(rows, nranks, tile_rows)] = image_rows[rows]
    return slab


def deinterleave(gathered: np.ndarray, height: int, nranks: int, tile_rows: int) -> np.ndarray:
    """[nranks, slab_rows, ...] gathered slabs -> [height, ...] image."""
    y = np.arange(height)
    owner = (y // tile_rows) % nranks
    return gathered[owner, slab_row_of(y, nranks, tile_rows)]
