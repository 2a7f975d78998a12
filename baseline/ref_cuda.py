#!/usr/bin/env python
"""The reference's OWN CUDA kernels re-hosted on the B200 -- the closest obtainable stand-in for "3D-SDN's PyTorch-0.4 /
Chainer / CuPy renderer on the same box" (SURVEY.md section 8c last row).

Chainer 4 / CuPy 4 / PyTorch 0.4 cannot be installed (no network; CUDA 9 never targeted sm_100), but the renderer's device
code is seven CUDA-C strings inside geometric/neural_renderer/rasterize.py.  At RUN time (nothing is committed) this
script reads that file from the git-ignored travelling copy baseline/_ref/geometric (or /root/reference), applies the same
string.Template substitutions the reference applies, wraps every body in CuPy's elementwise launch semantics
(one element per thread, raw arrays as plain pointers, 128 threads per block), compiles with NVRTC
for sm_100 (default --fmad=true, as CuPy does) and drives the kernels from torch tensors in the order
Rasterize.forward_gpu / backward_gpu issue them (rasterize.py:464-510, 846-886).

Measured workload: what one training step of 3D-SDN renders (derender3d/models/__init__.py:161-224, mode full): for each
object of a batch, ONE silhouette rasterisation at render_size 384 (768x768 internal, anti-aliasing) of its ShapeNet
template (fill_back: F = 2 nf), forward + backward with N(0,1) upstream gradients, objects one after the other --
  kernels : the reference's kernels only (K1 visibility [unsafe per-face kernel, NEURAL_RENDERER_UNSAFE=1 as scripts/env.sh
            sets it], alpha map, K6 backward), everything resident;
  bridge  : the same plus the host round trips of derender3d/models/renderer.py:134-150,183-211 re-enacted with the same
            byte counts (vertices / faces device->host->device, images device->host->device, gradients back), each a
            synchronous copy as `.cpu().numpy()` / `chainer.cuda.to_gpu` are.
The camera / gather / pooling glue around the kernels (Chainer ops in the reference) is done by small torch ops here; it is
not what is being compared.  Prints one JSON line.  Run under a timeout: the unsafe kernel spins on per-pixel locks."""
import argparse
import json
import math
import os
import re
import string
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402

KERNELS = [     # name, CuPy parameter list (rasterize.py:108-119, 242-246, 281-286, 373-379, 525-529, 757-761, 802-806)
    ('k1u', 'float* faces, int* face_index_map, float* weight_map, float* depth_map, float* face_inv_map, int* lock'),
    ('k1a', 'float* faces, float* faces_inv'),
    ('k1b', 'float* faces, float* faces_inv, int* face_index_map, float* weight_map, float* depth_map, float* face_inv_map'),
    ('k2', 'float* faces, float* textures, int* face_index_map, float* weight_map, float* depth_map, float* rgb_map, '
           'int* sampling_index_map, float* sampling_weight_map'),
    ('k6', 'float* faces, int* face_index_map, float* rgb_map, float* alpha_map, float* grad_rgb_map, float* grad_alpha_map, '
           'float* grad_faces'),
    ('k7', 'int* face_index_map, float* sampling_weight_map, int* sampling_index_map, float* grad_rgb_map, float* grad_textures'),
    ('k8', 'float* faces, float* depth_map, int* face_index_map, float* face_inv_map, float* weight_map, float* grad_depth_map, '
           'float* grad_faces'),
]


def find_rasterize_py():
    for root in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        p = os.path.join(root, 'geometric', 'neural_renderer', 'rasterize.py')
        if os.path.isfile(p):
            return p
    return None


def generate_source(path, cfg):
    """cfg: the values the reference substitutes (rasterize.py:226-233 etc.): image_size, num_faces, texture_size, near,
    far, eps, return_rgb / alpha / depth."""
    text = open(path).read()
    bodies = re.findall(r"string\.Template\('''(.*?)'''\)", text, flags=re.S)
    assert len(bodies) == len(KERNELS), 'expected %d kernel bodies, found %d' % (len(KERNELS), len(bodies))
    src = ['// generated at run time from the reference\'s rasterize.py -- never committed\n']
    for (name, params), body in zip(KERNELS, bodies):
        body = string.Template(body).substitute(**cfg)
        # one element per thread (the grid always covers n), so the bodies' `return` ends the element like CuPy's loop
        # does; a top-level `continue` leaves the do-while, i.e. also ends the element
        src.append('extern "C" __global__ void ref_%s(const int n, %s) {\n'
                   '  const int i = blockIdx.x * blockDim.x + threadIdx.x;\n'
                   '  if (i >= n) return;\n'
                   '  do {\n%s\n  } while (0);\n}\n' % (name, params, body))
    return '\n'.join(src)


class Module(object):
    def __init__(self, source, arch='sm_100'):
        from cuda.bindings import driver, nvrtc
        self.driver = driver
        err, prog = nvrtc.nvrtcCreateProgram(source.encode(), b'ref_kernels.cu', 0, [], [])
        assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS, err
        opts = [b'--gpu-architecture=' + arch.encode(), b'--std=c++14']
        err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
        if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
            _, n = nvrtc.nvrtcGetProgramLogSize(prog)
            log = b' ' * n
            nvrtc.nvrtcGetProgramLog(prog, log)
            raise RuntimeError('NVRTC: ' + log.decode(errors='replace')[-3000:])
        _, n = nvrtc.nvrtcGetCUBINSize(prog)
        cubin = b' ' * n
        nvrtc.nvrtcGetCUBIN(prog, cubin)
        torch.cuda.init()
        torch.zeros(1, device='cuda')                       # primary context current
        err, self.module = driver.cuModuleLoadData(cubin)
        assert err == driver.CUresult.CUDA_SUCCESS, err
        self.fn = {}
        for name, _ in KERNELS:
            err, f = driver.cuModuleGetFunction(self.module, ('ref_' + name).encode())
            assert err == driver.CUresult.CUDA_SUCCESS, (name, err)
            self.fn[name] = f

    def launch(self, name, n, *tensors):
        """CuPy elementwise launch: 128 threads per block, one element per thread (grid capped like CuPy's)."""
        import ctypes
        driver = self.driver
        args = [ctypes.c_int(n)] + [ctypes.c_void_p(t.data_ptr()) for t in tensors]
        arr = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
        blocks = max(1, min((n + 127) // 128, 2 ** 31 - 1))
        err, = driver.cuLaunchKernel(self.fn[name], blocks, 1, 1, 128, 1, 1, 0, torch.cuda.current_stream().cuda_stream,
                                     ctypes.addressof(arr), 0)
        assert err == driver.CUresult.CUDA_SUCCESS, (name, err)


def faces_ndc(vertices, faces, angle_deg):
    """derender3d renderer glue in torch (renderer.py:243-262 + look / perspective / vertices_to_faces / fill_back)."""
    v = vertices * torch.tensor([-1.0, 1.0, 1.0], device=vertices.device)         # renderer.py:243
    f = torch.cat([faces, faces.flip(-1)], dim=0)                                # fill_back, neural_renderer/renderer.py:41
    # look(eye 0, direction (0,0,-1), up (0,1,0)): rows x = (-1,0,0), y = (0,1,0), z = (0,0,-1), each 0.99999 long
    # (cf.normalize adds 1e-5 to the norm)
    x, y, z = -v[..., 0] * 0.99999, v[..., 1] * 0.99999, -v[..., 2] * 0.99999
    w = math.tan(angle_deg * 3.1416 / 180.0)
    ndc = torch.stack([x / z / w, y / z / w, z], dim=-1)
    return ndc[f.long()].contiguous().float()               # [F, 3, 3]


def silhouette_fwd_bwd(mod, fn, S, grad_alpha_ss, unsafe=True):
    """Rasterize.forward_gpu + backward_gpu for return_alpha only (rasterize.py:464-510, 846-886) on [1, F, 3, 3] faces."""
    dev = fn.device
    F_ = fn.shape[0]
    face_index = torch.full((S, S), -1, device=dev, dtype=torch.int32)
    weight = torch.zeros(S, S, 3, device=dev)
    depth = torch.full((S, S), 100.0, device=dev)
    dummy = torch.zeros(1, device=dev)
    if unsafe:
        lock = torch.zeros(S, S, device=dev, dtype=torch.int32)
        mod.launch('k1u', F_, fn, face_index, weight, depth, dummy, lock)
    else:
        inv = torch.zeros(F_, 3, 3, device=dev)
        mod.launch('k1a', F_, fn, inv)
        mod.launch('k1b', S * S, fn, inv, face_index, weight, depth, dummy)
    alpha = (face_index >= 0).float()                        # forward_alpha_map_gpu :437-446
    grad_faces = torch.zeros(F_, 3, 3, device=dev)
    mod.launch('k6', F_, fn, face_index, dummy, alpha, dummy, grad_alpha_ss, grad_faces)
    return alpha, grad_faces


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--objects', type=int, default=16)
    ap.add_argument('--render-size', type=int, default=384)
    ap.add_argument('--safe', action='store_true', help='K1a + K1b (O(S^2 F) per object) instead of the per-face kernel')
    ap.add_argument('--check', action='store_true', help='compare one object against this repo\'s renderer')
    args = ap.parse_args()
    path = find_rasterize_py()
    if path is None:
        print(json.dumps({'unavailable': 'rasterize.py of the reference is not reachable (baseline/_ref not staged)'}))
        return
    from derender3d.bench_setup import template_meshes
    from sdn_b200 import workloads as wl
    dev = torch.device('cuda:0')
    objs, desc = template_meshes()
    S = args.render_size * 2
    angle = wl.vkitti_viewing_angle(args.render_size, 500.0)
    gen = torch.Generator().manual_seed(3)
    mods = {}
    scenes = []
    for i in range(args.objects):
        o = objs[i % len(objs)]
        yaw = float(torch.rand(1, generator=gen)) * 6.283
        c, s_ = math.cos(yaw), math.sin(yaw)
        R = torch.tensor([[c, 0, s_], [0, 1, 0], [-s_, 0, c]])
        v = (o.vertices * torch.tensor([4.0, 1.6, 1.8])) @ R.T + torch.tensor([0.0, 0.0, -(5.5 + 3.0 * float(torch.rand(1, generator=gen)))])
        scenes.append((v.to(dev), o.faces.to(dev), torch.randn(args.render_size, args.render_size, generator=gen)))
        F_ = 2 * o.faces.shape[0]
        if F_ not in mods:      # the reference recompiles per (image_size, num_faces): CuPy caches by source text
            cfg = dict(image_size=S, num_faces=F_, texture_size=2, near=0.1, far=100, eps=1e-4, return_rgb=0, return_alpha=1,
                       return_depth=0)
            mods[F_] = Module(generate_source(path, cfg))

    def one(v, f, g, bridge):
        if bridge:      # renderer.py:134-150: tensors -> numpy -> chainer (to_gpu)
            v = torch.from_numpy(v.cpu().numpy()).to(v.device)
            f = torch.from_numpy(f.cpu().numpy()).to(f.device)
        fn = faces_ndc(v, f, angle)
        ga = (g.to(fn.device).flip(0) / 4.0).repeat_interleave(2, 0).repeat_interleave(2, 1).contiguous()   # pool + flip backward
        alpha, gf = silhouette_fwd_bwd(mods[fn.shape[0]], fn, S, ga, unsafe=not args.safe)
        img = torch.nn.functional.avg_pool2d(alpha.flip(0)[None, None], 2)[0, 0]
        if bridge:      # renderer.py:183-186 images -> host -> device; :195-211 gradients -> host -> device
            img = torch.from_numpy(img.cpu().numpy()).to(img.device)
            gf = torch.from_numpy(gf.cpu().numpy()).to(gf.device)
        return img, gf

    out = {'what': 'reference CUDA kernels (rasterize.py bodies through NVRTC, sm_100) on the B200: per-object silhouette '
                   'render fwd+bwd at %dx%d (%dx%d internal), %s, %s visibility kernel'
                   % (args.render_size, args.render_size, S, S, desc, 'safe (K1a+K1b)' if args.safe else 'unsafe per-face (K1u)')}
    for bridge in (False, True):
        one(*scenes[0], bridge)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for sc in scenes:
            img, gf = one(*sc, bridge)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out['bridge' if bridge else 'kernels'] = {'objects_per_s': round(len(scenes) / dt, 2), 'ms_per_object': round(dt / len(scenes) * 1e3, 2)}
    out['coverage'] = round(float((img > 0).float().mean()), 3)
    if args.check:
        from derender3d.models.renderer import Renderer, RenderType
        # this repo on the same scenes: ONE batched call (ragged topologies padded with degenerate triangles)
        nvm, nfm = max(v.shape[0] for v, _, _ in scenes), max(f.shape[0] for _, f, _ in scenes)
        vb = torch.zeros(len(scenes), nvm, 3, device=dev)
        fb = torch.zeros(len(scenes), nfm, 3, device=dev, dtype=torch.int32)
        for i, (v, f, g) in enumerate(scenes):
            vb[i, :v.shape[0]] = v
            vb[i, v.shape[0]:] = v[0]
            fb[i, :f.shape[0]] = f
        gb = torch.stack([g for _, _, g in scenes]).to(dev)[:, None]
        rr = Renderer(image_size=args.render_size, viewing_angle=angle)
        vbg = vb.requires_grad_(True)

        def ours():
            o = rr(vbg, fb, render_type=RenderType.Silhouette)
            o.backward(gb)
            vbg.grad = None
        for _ in range(3):
            ours()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ours()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        out['this_repo_same_scenes'] = {'objects_per_s': round(len(scenes) / dt, 1), 'ms_per_object': round(dt / len(scenes) * 1e3, 3),
                                        'how': 'one batched Renderer call, silhouette fwd+bwd, eager'}
        v, f, g = scenes[0]
        vg = v.clone().requires_grad_(True)
        mine = Renderer(image_size=args.render_size, viewing_angle=angle)(vg[None], f[None].int(), render_type=RenderType.Silhouette)
        img, gf = one(v, f, g, False)
        out['check'] = {'pixels_differing_from_this_repo': int((mine[0, 0] != img).sum()), 'pixels': int(img.numel())}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
