"""Frame sharding across GPUs and the validation gather.

The hot path has no collective: every frame is independent, so rank r of W simply owns a
contiguous block of frames and runs the single-GPU front end on it.  ``gather_outputs`` exists
for validation only -- it reassembles the per-rank outputs into what one GPU would have produced
for the whole batch, with the reference's size-then-payload protocol
(det3d/torchie/trainer/utils.py:114-153) but on typed tensors instead of pickled bytes.
Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_frames, world_size, rank):
    """Contiguous block [lo, hi) of frames owned by ``rank``."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_frames, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _all_gather_rows(t, group=None):
    """all_gather of tensors whose first dimension differs per rank: sizes first, then payloads
    padded to the largest size."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:s] for b, s in zip(bufs, sizes)]


def gather_outputs(local, n_local_frames, group=None):
    """local: dict with 'coordinates' [M,4] (b,z,y,x; b local), 'num_points' [M], 'num_voxels' [B_local]
    and optionally 'features' [M,C], 'voxels' [M,T,C], 'canvas' [B_local,C,ny,nx].
    Returns the dict every rank would hold had one GPU processed all frames (batch index global)."""
    world = dist.get_world_size(group)
    dev = local["coordinates"].device
    nf = torch.tensor([n_local_frames], dtype=torch.int64, device=dev)
    all_nf = [torch.zeros_like(nf) for _ in range(world)]
    dist.all_gather(all_nf, nf, group=group)
    frame_base = [0]
    for x in all_nf:
        frame_base.append(frame_base[-1] + int(x.item()))
    out = {}
    coords = _all_gather_rows(local["coordinates"], group)
    fixed = []
    for r, c in enumerate(coords):
        c = c.clone()
        if c.numel():
            c[:, 0] += frame_base[r]
        fixed.append(c)
    out["coordinates"] = torch.cat(fixed, dim=0)
    for key in ("num_points", "num_voxels", "features", "voxels", "canvas"):
        if key in local and local[key] is not None:
            out[key] = torch.cat(_all_gather_rows(local[key], group), dim=0)
    return out
