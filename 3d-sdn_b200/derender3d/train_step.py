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


def _partial(f, mask):
    """main.py:97-112: apply f on the selected rows only; 0 when none is selected."""
    def g(*args, **kw):
        index = torch.nonzero(mask)
        if index.numel():
            index = index.squeeze(dim=1)
            return f(*[a[index] for a in args], **kw)
        return torch.zeros((), device=mask.device)
    return g


def geometric_losses(blob, batch, mode, mask_weight=MASK_WEIGHT, ffd_coeff_reg=FFD_COEFF_REG):
    """blob: output of Derenderer3d.forward; batch: dict with targets [B] (TargetType bits), thetas [B,1],
    translation2ds [B,2], log_scales [B,3], log_depths [B,1], masks / ignores [B,1,h,w].  -> dict of scalar losses."""
    losses = {}
    targets = batch['targets']
    if mode & TargetType.geometry:
        mse = _partial(F.mse_loss, targets & TargetType.pretrain)
        theta_deltas = torch.cat([torch.cos(batch['thetas']), torch.sin(batch['thetas'])], dim=1)
        losses['theta_delta_loss'] = mse(blob['_theta_deltas'], theta_deltas)
        losses['translation2d_loss'] = mse(blob['_translation2ds'], batch['translation2ds'])
        losses['scale_loss'] = mse(blob['_log_scales'], batch['log_scales'])
        losses['depth_loss'] = mse(blob['_log_depths'], batch['log_depths'])
    if mode & TargetType.reproject:
        sel = targets & TargetType.finetune
        mean = _partial(torch.mean, sel)
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
