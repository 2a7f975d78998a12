"""Analysis-by-synthesis refinement of the de-rendered latents (reference: geometric/scripts/main.py:405-459):
`num_opts` Adam steps (lr 3e-2) on theta / 2-D translation / log-scale / FFD coefficients through Derenderer3d.render
with a silhouette MSE + 100 * mean(ffd^2).  Every iteration is one batched render forward + backward -- a pure
rasteriser-throughput workload; nothing touches the host inside the loop except the optional loss
read-out, so with graph=True iteration 0 runs eagerly (lazy initialisation, Adam state) and iterations 1..n-1 are
replays of ONE captured CUDA graph (render forward, loss, backward, Adam)."""
import torch

from derender3d.train_step import pad_like

OPT_KEYS = ('_theta_deltas', '_translation2ds', '_log_scales', '_ffd_coeffs')       # main.py:430-435


def refine(model, blob, masks, ignores=None, num_opts=32, lr=3e-2, ffd_weight=100.0, log=None, graph=False):
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
    graph = bool(graph) and num_opts > 1 and params[0].is_cuda
    opt = torch.optim.Adam(params, lr=lr, capturable=graph)

    def iteration(zero=True):
        if zero:
            opt.zero_grad(set_to_none=True)
        out = model.render(blob)
        _masks = out['_masks']
        loss = torch.nn.functional.mse_loss(_masks, pad_like(masks, _masks), reduction='none') + \
            ffd_weight * torch.mean(blob['_ffd_coeffs'] ** 2)
        if ignores is not None:
            loss = loss * (1 - pad_like(ignores, _masks, mode='replicate'))
        loss = torch.mean(loss)
        loss.backward()
        opt.step()
        return loss.detach(), out

    history = []

    def record(it, loss):
        history.append(loss.clone())
        if log is not None:
            log('Optimizing {:d}/{:d}: loss={:.4e}'.format(it + 1, num_opts, float(loss)))

    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            loss, out = iteration()
        torch.cuda.current_stream().wait_stream(side)
        record(0, loss)
        opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss, out = iteration(zero=False)
        for it in range(1, num_opts):
            g.replay()
            record(it, loss)
    else:
        for it in range(num_opts):
            loss, out = iteration()
            record(it, loss)
    model.train(was_training)
    model._force_no_sample = was_force
    blob.update({k: v.detach() for k, v in out.items() if isinstance(v, torch.Tensor)})
    for k in OPT_KEYS:
        blob[k] = blob[k].detach()
    return blob, torch.stack(history)
