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
    """All-reduce (mean) of all gradients through one flat bucket: a single collective launch per step.

    bucket_views=True makes every `p.grad` a VIEW of the flat buffer (autograd accumulates into an existing .grad in
    place), so the step is: `zero()` (one memset) -> backward -> `__call__()` (one all-reduce + one scale, no copies).
    Keep the views alive: clear gradients with `zero()` / `optimizer.zero_grad(set_to_none=False)`, never set_to_none.
    Everything here is capturable in a CUDA graph (NCCL collectives included)."""

    def __init__(self, model, group=None, bucket_views=False):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.numel = n
        self.views, off = [], 0
        for p in self.params:
            chunk = self.flat[off:off + p.numel()]
            if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                # channels-last parameter: give its gradient view the same strides (autograd's layout contract), so
                # AccumulateGrad adds in place instead of re-striding through a temporary
                n, c, h, w = p.shape
                self.views.append(chunk.view(n, h, w, c).permute(0, 3, 1, 2))
            else:
                self.views.append(chunk.view_as(p))
            off += p.numel()
        self.bucket_views = bool(bucket_views)
        if self.bucket_views:
            self.attach()

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        if self.bucket_views and not self._attached():
            self.attach()       # something (zero_grad(set_to_none=True), a layout change) dropped the views: re-attach
        self.flat.zero_()

    def _attached(self):
        return all(p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(self.params, self.views))

    def __call__(self, model=None):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        in_place = self.bucket_views and self._attached()
        if not in_place:
            have = [(v, p.grad) for p, v in zip(self.params, self.views) if p.grad is not None]
            for p, v in zip(self.params, self.views):
                if p.grad is None:
                    v.zero_()
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(world)
        if not in_place:
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.empty_like(p)
            torch._foreach_copy_([p.grad for p in self.params], self.views)
