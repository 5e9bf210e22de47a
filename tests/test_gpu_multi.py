"""Shard equivalence on real GPUs: frames split over W ranks (one process per GPU, NCCL) and
gathered back equal the single-GPU result bit for bit.  Runs with W = 1 always, W = 2 when the
box has two GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_rank(rank, world, port, ok):
    import torch.distributed as dist
    from partner_b200 import PolarFrontEnd, synth
    from partner_b200.sharding import gather_outputs, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        g = synth.GRIDS["NUSC-PILLAR"]
        frames = [synth.nusc_frame(40 + f, nsweeps=2) for f in range(4)]
        fe = PolarFrontEnd(g["voxel_size"], g["range"], g["max_points"], 20000, device="cuda:%d" % rank)
        lo, hi = shard_range(len(frames), world, rank)
        loc = fe(frames[lo:hi])
        dev = torch.device("cuda", rank)
        local = {k: torch.from_numpy(v).to(dev) for k, v in
                 dict(coordinates=loc["coordinates"], num_points=loc["num_points"], num_voxels=loc["num_voxels"],
                      features=loc["features"], canvas=loc["canvas"]).items()}
        got = gather_outputs(local, hi - lo)
        ref = fe(frames)
        for k in ("coordinates", "num_points", "num_voxels"):       # integer outputs: bit-exact
            assert np.array_equal(got[k].cpu().numpy(), ref[k]), k
        for k in ("features", "canvas"):                            # fp32 sums in unspecified order
            a, b = got[k].cpu().numpy(), ref[k]
            assert a.shape == b.shape, k
            assert np.allclose(a, b, rtol=1e-5, atol=1e-5 * float(np.abs(b).max())), k
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_sharded_equals_single_gpu(world):
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_run_rank, args=(world, port, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
