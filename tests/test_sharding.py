"""Host-side multi-GPU logic on CPU: frame sharding and the typed gather protocol over gloo
(world_size 2), mirroring how the N>1 bench / validation path reassembles per-rank outputs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from partner_b200.sharding import gather_outputs, shard_range


def test_shard_range_partitions_every_frame_once():
    for n in (0, 1, 7, 8, 64, 65):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = shard_range(n, w, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, w, r)[1] - shard_range(n, w, r)[0] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _fake_frame_outputs(frame, c=7, t=4):
    """Deterministic per-frame outputs standing in for the front end (no GPU here)."""
    rng = np.random.default_rng(100 + frame)
    m = int(rng.integers(0, 40)) if frame != 2 else 0          # frame 2 is empty
    return dict(coors=rng.integers(0, 512, (m, 3)).astype(np.int32), num=rng.integers(1, t + 1, m).astype(np.int32),
                feats=rng.normal(size=(m, c)).astype(np.float32), canvas=rng.normal(size=(1, c, 4, 4)).astype(np.float32))


def _collate(frames, first=0):
    outs = [_fake_frame_outputs(f) for f in frames]
    coords = np.concatenate([np.pad(o["coors"], ((0, 0), (1, 0)), constant_values=i) for i, o in enumerate(outs)]
                            or [np.zeros((0, 4), np.int32)])
    return dict(coordinates=torch.from_numpy(coords.astype(np.int32)),
                num_points=torch.from_numpy(np.concatenate([o["num"] for o in outs] or [np.zeros(0, np.int32)])),
                num_voxels=torch.tensor([o["coors"].shape[0] for o in outs], dtype=torch.int64),
                features=torch.from_numpy(np.concatenate([o["feats"] for o in outs] or [np.zeros((0, 7), np.float32)])),
                canvas=torch.from_numpy(np.concatenate([o["canvas"] for o in outs] or [np.zeros((0, 7, 4, 4), np.float32)])))


def _worker(rank, world, port, n_frames, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_frames, world, rank)
        local = _collate(list(range(lo, hi)))
        got = gather_outputs(local, hi - lo)
        ref = _collate(list(range(n_frames)))
        for k in ref:
            assert torch.equal(got[k], ref[k]), k
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8])
def test_gather_outputs_gloo_world2_equals_single_rank(n_frames):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_worker, args=(2, port, n_frames, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
