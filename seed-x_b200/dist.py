"""Multi-GPU plumbing: requests are independent, so ranks are replicas.  The only collective on the data path is
the gather of the finished images; timing uses a MAX over ranks.  Works with NCCL (GPU) and gloo (CPU tests)."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_requests(n_requests, r=None, w=None):
    """indices of the requests rank r serves: r, r+w, r+2w, ... (one request set per rank, cfg4/cfg5 of BASELINE.json)"""
    r = rank() if r is None else r
    w = world() if w is None else w
    return list(range(r, n_requests, w))


def gather_images(u8, buf=None):
    """all ranks end up with every rank's [B,H,W,3] uint8 images, stacked [world, B, H, W, 3] (the single data-path collective)"""
    w = world()
    if w == 1:
        return u8.unsqueeze(0)
    if buf is None or buf.shape[1:] != u8.shape:
        buf = torch.empty((w,) + tuple(u8.shape), device=u8.device, dtype=u8.dtype)
    dist.all_gather_into_tensor(buf.view((w * u8.shape[0],) + tuple(u8.shape[1:])), u8.contiguous())   # concatenated form (gloo + NCCL)
    return buf


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if world() > 1:
        dist.barrier()
