"""One optimisation step of the geometric branch (reference: geometric/scripts/main.py:96-154 `BaseNet.step_batch`,
bulb/net.py:164-167 `_optimize`): masked MSE on the pose heads, silhouette reprojection loss, REINFORCE on the mesh
class, FFD regulariser.  main.py itself is the caller and stays out of scope; this module lets the benchmark and the
tests run the same arithmetic without absl / bulb / tensorboardX."""
import torch
import torch.nn.functional as F

from derender3d import TargetType

MASK_WEIGHT = 0.1       # main.py:41
FFD_COEFF_REG = 1.0     # main.py:42


def pad_like(image, like, mode='constant', value=0):
    """derender3d/datasets.py:29-33."""
    p2, p3 = like.shape[2] - image.shape[2], like.shape[3] - image.shape[3]
    if mode == 'constant':
        return F.pad(image, (p3 // 2, p3 // 2, p2 // 2, p2 // 2), mode=mode, value=value)
    return F.pad(image, (p3 // 2, p3 // 2, p2 // 2, p2 // 2), mode=mode)


def _masked_mse(a, b, mask):
    """main.py:97-112 `partial(F.mse_loss, mask)`: mean squared error over the selected rows, 0 when none is selected --
    written without torch.nonzero so that nothing syncs with the host (the step can then be captured in a CUDA graph)."""
    w = (mask != 0).to(a.dtype)
    num = (((a - b) ** 2).reshape(a.shape[0], -1).sum(dim=1) * w).sum()
    den = w.sum() * (a.numel() // max(a.shape[0], 1))
    return torch.where(den > 0, num / den.clamp_min(1.0), torch.zeros((), device=a.device, dtype=a.dtype))


def _masked_mean(x, mask):
    """`partial(torch.mean, mask)` of a [B] vector."""
    w = (mask != 0).to(x.dtype)
    den = w.sum()
    return torch.where(den > 0, (x * w).sum() / den.clamp_min(1.0), torch.zeros((), device=x.device, dtype=x.dtype))


def geometric_losses(blob, batch, mode, mask_weight=MASK_WEIGHT, ffd_coeff_reg=FFD_COEFF_REG):
    """blob: output of Derenderer3d.forward; batch: dict with targets [B] (TargetType bits), thetas [B,1],
    translation2ds [B,2], log_scales [B,3], log_depths [B,1], masks / ignores [B,1,h,w].  -> dict of scalar losses."""
    losses = {}
    targets = batch['targets']
    if mode & TargetType.geometry:
        pre = targets & TargetType.pretrain
        mse = lambda a, b: _masked_mse(a, b, pre)      # noqa: E731
        theta_deltas = torch.cat([torch.cos(batch['thetas']), torch.sin(batch['thetas'])], dim=1)
        losses['theta_delta_loss'] = mse(blob['_theta_deltas'], theta_deltas)
        losses['translation2d_loss'] = mse(blob['_translation2ds'], batch['translation2ds'])
        losses['scale_loss'] = mse(blob['_log_scales'], batch['log_scales'])
        losses['depth_loss'] = mse(blob['_log_depths'], batch['log_depths'])
    if mode & TargetType.reproject:
        sel = targets & TargetType.finetune
        mean = lambda x: _masked_mean(x, sel)          # noqa: E731
        masks = pad_like(batch['masks'], blob['_masks'])
        ignores = pad_like(batch['ignores'], blob['_masks'], mode='replicate')
        mask_losses = (1 - ignores) * F.mse_loss(blob['_masks'], masks, reduction='none')
        mask_losses = mask_weight * mask_losses.mean(dim=3).mean(dim=2).mean(dim=1)
        losses['class_reward'] = mean(blob['_class_log_probs'] * mask_losses.detach())       # REINFORCE, main.py:149
        losses['mask_loss'] = mean(mask_losses)
        losses['ffd_coeff_reg'] = ffd_coeff_reg * torch.mean(blob['_ffd_coeffs'] ** 2)
    return losses


def train_step(model, optimizer, batch, mode, grad_sync=None):
    """bulb/net.py:121,164-167: loss = sum of the losses; zero_grad; backward; [all-reduce]; step."""
    blob = model(batch['images'], batch['roi_norms'], batch['focals'])
    losses = geometric_losses(blob, batch, mode)
    loss = sum(losses.values())
    optimizer.zero_grad(set_to_none=False)
    loss.backward()
    if grad_sync is not None:
        grad_sync(model)
    optimizer.step()
    return loss.detach(), losses, blob


class GraphedTrainStep(object):
    """The whole optimisation step -- encoder, pose algebra, class sampling, fused pre-raster stage, rasterisation,
    losses, backward, Adam -- captured ONCE as a CUDA graph and replayed per batch (one launch per step instead of
    ~1500).  Possible because nothing on the path syncs with the host: the mesh class is a device-side index
    (sdn_objects), the losses use masks instead of torch.nonzero, shapes and the workspace are static.

    `optimizer` must be built with capturable=True (Adam keeps its step counter on the device).  Inputs are copied
    into static buffers; outputs (`loss`, `losses`, `blob`) are static tensors overwritten by each replay.

    The warm-up steps that lazy initialisation needs (mesh bank, camera cache, cuDNN plans, Adam state, NCCL
    communicator) do NOT leak into the training trajectory: parameters, buffers (BatchNorm running statistics) and the
    optimizer state are snapshotted before and restored after them, so replay k equals eager step k.

    Multi-GPU (BASELINE config 4): pass `grad_sync` = sdn_b200.dist.FlatGradSync(model, bucket_views=True).  The
    gradients are then views of one flat fp32 bucket and the captured step is
    zero bucket -> forward -> backward -> ONE all-reduce (NCCL, captured in the graph) -> Adam; every rank replays the
    same graph, so the replicas stay bit-identical (checked by `replicas_in_sync`)."""

    def __init__(self, model, optimizer, example_batch, mode, warmup=3, autocast_dtype=None, grad_sync=None):
        import copy
        self.model, self.optimizer, self.mode = model, optimizer, mode
        self.batch = {k: v.clone() for k, v in example_batch.items()}
        self.autocast_dtype = autocast_dtype
        self.grad_sync = grad_sync
        snap_model = {k: v.detach().clone() for k, v in model.state_dict().items()}
        snap_opt = copy.deepcopy(optimizer.state_dict())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):     # lazy initialisation; undone below
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(snap_model[k])
        if snap_opt['state']:
            optimizer.load_state_dict(snap_opt)
        else:                                   # first use: forget the moments and step count of the warm-up
            with torch.no_grad():
                for st in optimizer.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
        if grad_sync is None:
            optimizer.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss, self.losses, self.blob = self._eager(zero=grad_sync is not None)

    def _eager(self, zero=True):
        with torch.autocast('cuda', dtype=self.autocast_dtype or torch.bfloat16, enabled=self.autocast_dtype is not None):
            blob = self.model(self.batch['images'], self.batch['roi_norms'], self.batch['focals'])
            losses = geometric_losses(blob, self.batch, self.mode)
            loss = sum(losses.values())
        if self.grad_sync is not None:
            self.grad_sync.zero()               # gradients are views of the flat bucket: one memset, views stay alive
        elif zero:
            self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync(self.model)
        self.optimizer.step()
        return loss.detach(), losses, blob

    def replicas_in_sync(self):
        """True when every rank holds bit-identical parameters (all-gather of a checksum per rank)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return True
        with torch.no_grad():
            flat = torch.cat([p.detach().reshape(-1).float() for p in self.model.parameters()])
            sig = torch.stack([flat.double().sum(), flat.double().abs().sum(), (flat.double() * torch.arange(
                flat.numel(), device=flat.device, dtype=torch.float64).remainder(97.0)).sum()])
            sigs = [torch.empty_like(sig) for _ in range(dist.get_world_size())]
            dist.all_gather(sigs, sig)
        return all(torch.equal(sigs[0], s) for s in sigs[1:])

    def step(self, batch=None):
        if batch is not None:
            for k, v in batch.items():
                self.batch[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss
