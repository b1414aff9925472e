"""World-size-2 `gloo` test of the N>1 host logic (runs on CPU): frames are sharded one per rank, every rank
packs its frame record and one all-gather per step hands every rank every frame, in frame order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

H, W, CAP = 8, 16, 40


def _fake_frame(frame_id):
    from sivo_b200.orb import KP_DTYPE
    rng = np.random.default_rng(100 + frame_id)
    nl, nr = int(rng.integers(1, CAP)), int(rng.integers(1, CAP))
    mk = lambda n: np.array([tuple(rng.normal(size=5).astype(np.float32)) + (int(rng.integers(0, 8)), -1) for _ in range(n)], KP_DTYPE)
    return dict(classes=rng.integers(0, 15, (H, W), dtype=np.uint8), confidence=rng.random((H, W)), entropy=rng.random((H, W)),
                kl=mk(nl), dl=rng.integers(0, 256, (nl, 32), dtype=np.uint8), kr=mk(nr), dr=rng.integers(0, 256, (nr, 32), dtype=np.uint8))


def _worker(rank, world, port, steps, q):
    from sivo_b200 import record
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hw = H * W
    nbytes = record.record_bytes(hw, CAP)
    ok = True
    for step in range(steps):
        fid = step * world + rank  # the sharding rule bench.py uses
        f = _fake_frame(fid)
        rec = np.zeros(nbytes, np.uint8)
        o = record.offsets(hw, CAP)
        if step % 2 == 0:  # the device-record layout written piecewise (value path: maps by SegNet, the rest by the extractors)
            record.pack_host_part(rec, hw, CAP, fid, f["kl"], f["dl"], f["kr"], f["dr"])
            rec[o["classes"]:o["classes"] + hw] = f["classes"].reshape(-1)
            rec[o["confidence"]:o["confidence"] + hw * 4] = f["confidence"].astype(np.float32).view(np.uint8).reshape(-1)
            rec[o["entropy"]:o["entropy"] + hw * 4] = f["entropy"].astype(np.float32).view(np.uint8).reshape(-1)
        else:              # the whole record from host results (N > 1 e2e path)
            record.pack_host(rec, hw, CAP, fid, f["classes"], f["confidence"], f["entropy"], f["kl"], f["dl"], f["kr"], f["dr"])
        # the header words the asynchronous extractors write on the device: frame id, n_left, n_right as int64 at bytes 0 / 8 / 16
        assert tuple(int(v) for v in rec[:24].view(np.int64)) == (fid, len(f["kl"]), len(f["kr"]))
        mine = torch.from_numpy(rec)
        allr = torch.empty(nbytes * world, dtype=torch.uint8)
        dist.all_gather_into_tensor(allr, mine)
        for r in range(world):
            u = record.unpack(allr[r * nbytes:(r + 1) * nbytes].numpy(), H, W, CAP)
            g = _fake_frame(step * world + r)
            ok &= u["frame_id"] == step * world + r
            ok &= np.array_equal(u["classes"], g["classes"]) and np.array_equal(u["entropy"], g["entropy"].astype(np.float32))
            ok &= np.array_equal(u["confidence"], g["confidence"].astype(np.float32))
            ok &= np.array_equal(u["kp_left"], g["kl"]) and np.array_equal(u["desc_right"], g["dr"])
            ok &= np.array_equal(u["kp_right"], g["kr"]) and np.array_equal(u["desc_left"], g["dl"])
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction
    ok &= float(t) == float(world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_frame_sharding_and_record_allgather_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 4, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, True), (1, True)]
