"""Multi-GPU plumbing of the render path.  The rasteriser shards by image with no exchange at all;
the only collective of a training step is ONE all-reduce over a single flat fp32 buffer holding every parameter
gradient (11.84 M elements = 47.4 MB for the Derenderer), replacing torch.nn.DataParallel's per-step broadcast /
scatter / gather / reduce-add (geometric/scripts/main.py:182).  One process per GPU (torchrun), NCCL on GPUs, gloo in
the CPU tests."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradSync(object):
    """All-reduce (mean) of all gradients through one flat bucket: a single collective launch per step."""

    def __init__(self, model, group=None):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.numel = n

    def __call__(self, model=None):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(world)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = torch.empty_like(p)
            p.grad.copy_(self.flat[off:off + n].view_as(p))
            off += n
