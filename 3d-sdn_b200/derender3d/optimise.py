"""Analysis-by-synthesis refinement of the de-rendered latents (reference: geometric/scripts/main.py:405-459):
`num_opts` Adam steps (lr 3e-2) on theta / 2-D translation / log-scale / FFD coefficients through Derenderer3d.render
with a silhouette MSE + 100 * mean(ffd^2).  Every iteration is one batched render forward + backward -- a pure
rasteriser-throughput workload; nothing touches the host inside the loop except the optional loss
read-out."""
import torch

from derender3d.train_step import pad_like

OPT_KEYS = ('_theta_deltas', '_translation2ds', '_log_scales', '_ffd_coeffs')       # main.py:430-435


def refine(model, blob, masks, ignores=None, num_opts=32, lr=3e-2, ffd_weight=100.0, log=None):
    """model: Derenderer3d; blob: encoder output + ROI/focal entries (detached copies are optimised);
    masks [B,1,h,w] target silhouettes, ignores [B,1,h,w] or None.  Returns (blob with refined latents, losses)."""
    was_training, was_force = model.training, model._force_no_sample
    model.train()                       # main.py:424-425: train-mode perspective transform, but no class sampling
    model._force_no_sample = True
    blob = dict(blob)
    for k, v in list(blob.items()):
        if isinstance(v, torch.Tensor):
            blob[k] = v.clone().detach()
    params = [blob[k].requires_grad_() for k in OPT_KEYS]
    opt = torch.optim.Adam(params, lr=lr)
    history = []
    for it in range(num_opts):
        opt.zero_grad()
        out = model.render(blob)
        _masks = out['_masks']
        loss = torch.nn.functional.mse_loss(_masks, pad_like(masks, _masks), reduction='none') + \
            ffd_weight * torch.mean(blob['_ffd_coeffs'] ** 2)
        if ignores is not None:
            loss = loss * (1 - pad_like(ignores, _masks, mode='replicate'))
        loss = torch.mean(loss)
        loss.backward()
        opt.step()
        history.append(loss.detach())
        if log is not None:
            log('Optimizing {:d}/{:d}: loss={:.4e}'.format(it + 1, num_opts, float(loss)))
    model.train(was_training)
    model._force_no_sample = was_force
    blob.update({k: v.detach() for k, v in out.items() if isinstance(v, torch.Tensor)})
    for k in OPT_KEYS:
        blob[k] = blob[k].detach()
    return blob, torch.stack(history)
