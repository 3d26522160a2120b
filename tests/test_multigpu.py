"""N > 1 on real GPUs: two ranks (one process per GPU) render their row tiles and
the frame is assembled on every rank -- by direct placement into peer-mapped frames
(default) and by the NCCL all-gather + de-interleave fallback; every rank must end
up with the full frame, bit-identical to the single-rank frame and the oracle, frame
after frame.
Skipped with fewer than two devices (the CPU/gloo twin is tests/test_multirank_cpu.py)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import clpathtracer_b200 as cl
from clpathtracer_b200 import scenes
from oracle import oracle_py as op

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = cl.lib()
scene = cl.build_kd_sah(*scenes.heightfield(40, False))
w, h = 250, 131                      # neither a multiple of the tile size
cam = cl.cam_matrix(cl.make_camera(**scenes.CANONICAL_CAMERA), h)
r = cl.Renderer(device=local)
r.set_meshes(scene)
r.set_camera_matrix(cam)
def fresh_id():
    """An NCCL unique id is good for one communicator: rank 0 makes one per CLDistInit."""
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = np.zeros(128, dtype=np.uint8)
        L.CLDistGetUniqueId(raw.ctypes.data)
        idbuf = torch.from_numpy(raw.copy())
    idbuf = idbuf.cuda()
    dist.broadcast(idbuf, 0)
    return idbuf.cpu().numpy().copy()

ok = True
refs = {seed: op.render(scene, cam, w, h, mode=1, depth=4, spp=5, seed=seed, flags=op.FLAG_JITTER, aov=False)["rgba"]
        for seed in (3, 4)}
for tile_rows, engine, direct in ((8, 1, 1), (4, 1, 1), (8, 2, 1), (8, 1, 0)):
    os.environ["CLPT_P2P"] = str(direct)
    raw = fresh_id()
    L.CLDistInit(rank, world, raw.ctypes.data, tile_rows)
    L.CLSetEngine(engine)
    r.create_image(w, h)
    got_direct = L.CLDistDirectPlacement()
    for seed in (3, 4):                # two frames back to back: no stale or torn rows
        r.set_params(mode=1, depth=4, spp=5, seed=seed, flags=cl.FLAG_JITTER)
        r.execute()
        img = r.read_image()
        same = np.array_equal(img.view(np.uint32), refs[seed].view(np.uint32)) and got_direct == direct
        print(f"rank {rank} tile_rows {tile_rows} engine {engine} direct {got_direct} seed {seed}: "
              f"{'ok' if same else 'MISMATCH'}", flush=True)
        ok = ok and same
    # progressive: spread by SAMPLE -- every rank renders the whole frame with its own sample indices into
    # its own fixed-point sums, nothing crosses GPUs per frame; the (collective) read-back adds the ranks'
    # sums and normalises.  Three frames of 2 spp on `world` ranks = samples 0 .. 6*world-1.
    r.set_params(mode=1, depth=4, spp=2, seed=9, flags=cl.FLAG_JITTER | cl.FLAG_ACCUMULATE)
    r.create_image(w, h)
    for _ in range(3):
        r.execute()
    mean = op.render(scene, cam, w, h, mode=1, depth=4, spp=6 * world, seed=9, aov=False,
                     flags=op.FLAG_JITTER | op.FLAG_ACCUMULATE)["rgba"]
    q = np.rint(np.clip(mean, 0.0, 1.0).astype(np.float32) * np.float32(255.0)).astype(np.uint8)
    same = np.array_equal(r.read_image().view(np.uint32), mean.view(np.uint32)) and \
        np.array_equal(r.read_image_rgba8(), q)
    print(f"rank {rank} tile_rows {tile_rows} engine {engine} direct {got_direct} progressive: "
          f"{'ok' if same else 'MISMATCH'}", flush=True)
    ok = ok and same
    L.CLDistShutdown()
os.environ.pop("CLPT_P2P", None)
r.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
'''


@pytest.mark.gpu
def test_two_rank_frame_assembly(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": str(ROOT)})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count(": ok") == 24, p.stdout
    print(p.stderr[-1500:])  # shown with -s: mapping diagnostics
