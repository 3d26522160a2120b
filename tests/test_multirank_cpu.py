"""The N > 1 path on CPU: two gloo ranks each render their row tiles (with the
oracle standing in for the device), exchange slabs with all_gather and
de-interleave; the result must be the single-rank frame.  This covers the
host-side sharding arithmetic that csrc/cuda/clstate.cu and the kernels
implement on the device (clpathtracer_b200/sharding.py states it)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tile_rows, out_dir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    import clpathtracer_b200 as cl
    from clpathtracer_b200 import scenes, sharding
    from oracle import oracle_py as op

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 64, 50
    scene = cl.build_kd(*scenes.heightfield(22, True))
    cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
    rows = sharding.rows_of_rank(h, rank, world, tile_rows)
    mine = np.zeros((h, w, 4), dtype=np.float32)
    # render only this rank's rows, tile by tile
    for y in rows:
        part = op.render(scene, cam, w, h, mode=1, depth=3, rows=(int(y), int(y) + 1), aov=False, threads=1)
        mine[y] = part["rgba"][y]
    slab = torch.from_numpy(sharding.to_slab(mine, h, rank, world, tile_rows))
    gathered = [torch.empty_like(slab) for _ in range(world)]
    dist.all_gather(gathered, slab)
    image = sharding.deinterleave(np.stack([g.numpy() for g in gathered]), h, world, tile_rows)
    if rank == 0:
        full = op.render(scene, cam, w, h, mode=1, depth=3, aov=False)["rgba"]
        np.save(Path(out_dir) / "ok.npy", np.array([np.array_equal(image, full)]))
    dist.barrier()
    dist.destroy_process_group()


def _progressive_worker(rank, world, port, out_dir):
    """Progressive frames spread by sample: every rank adds its samples of three frames to its own
    fixed-point sums (the oracle standing in for the device), an all-reduce adds the ranks' sums."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    import clpathtracer_b200 as cl
    from clpathtracer_b200 import scenes, sharding
    from oracle import oracle_py as op

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, spp, frames = 48, 32, 2, 3
    scene = cl.build_kd(*scenes.heightfield(22, True))
    cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
    kw = dict(mode=1, depth=3, seed=4, aov=False, threads=1, flags=op.FLAG_JITTER | op.FLAG_ACCUMULATE)
    acc = op.new_accumulator(w, h)
    for k in range(frames):
        op.render(scene, cam, w, h, spp=spp, sample_base=sharding.first_sample_of_rank(k, rank, world, spp),
                  accumulate_into=acc, **kw)
    total = torch.from_numpy(acc.view(np.int64).copy())   # (gloo has no uint64; the sums are far below 2^63)
    dist.all_reduce(total)
    if rank == 0:
        one = op.render(scene, cam, w, h, spp=spp * world * frames, **kw)
        np.save(Path(out_dir) / "ok.npy", np.array([np.array_equal(total.numpy().view(np.uint64), one["accum"])]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_progressive_by_sample(tmp_path, world, clpt, oracle):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_progressive_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "ok.npy")[0]


@pytest.mark.parametrize("world,tile_rows", [(2, 8), (3, 4)])
def test_gloo_row_tile_gather(tmp_path, world, tile_rows, clpt, oracle):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(world, port, tile_rows, str(tmp_path)), nprocs=world, join=True)
    assert np.load(tmp_path / "ok.npy")[0]


def test_sharding_arithmetic():
    from clpathtracer_b200 import sharding

    for h, n, tr in [(1080, 8, 8), (1080, 3, 16), (50, 2, 8), (7, 4, 4), (2160, 8, 8)]:
        seen = np.zeros(h, dtype=int)
        for r in range(n):
            rows = sharding.rows_of_rank(h, r, n, tr)
            seen[rows] += 1
            assert sharding.slab_row_of(rows, n, tr).max(initial=-1) < sharding.slab_rows(h, n, tr)
            assert len(np.unique(sharding.slab_row_of(rows, n, tr))) == len(rows)
        assert (seen == 1).all()
        img = np.arange(h * 3, dtype=np.float32).reshape(h, 3)
        slabs = np.stack([sharding.to_slab(img, h, r, n, tr) for r in range(n)])
        assert np.array_equal(sharding.deinterleave(slabs, h, n, tr), img)
