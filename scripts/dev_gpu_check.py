import sys, os, time
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'3d-sdn_b200'))
import numpy as np, torch
from oracle import oracle as O
import neural_renderer as nr
import sdn_b200
from sdn_b200 import workloads as wl
from derender3d.models.renderer import Renderer, RenderType
dev='cuda'
print('lib version', sdn_b200.lib().sdn_version())
def rel(a,b):
    a=a.detach().cpu().double(); b=b.detach().cpu().double()
    return ((a-b).norm()/(b.norm()+1e-30)).item()
# --- cfg-1 cube
v,f = wl.cube_mesh()
V=torch.tensor(v)[None]; Fc=torch.tensor(f)[None]
eye = nr.get_points_from_angles(2.732,20,30)
for aa in (True, False):
  for size in (64, (48,80)):
    ro = O.NRRenderer(); ro.image_size=size; ro.eye=eye; ro.anti_aliasing=aa
    rg = nr.Renderer(); rg.image_size=size; rg.eye=eye; rg.anti_aliasing=aa
    tex = torch.rand(1,12,2,2,2,3, generator=torch.Generator().manual_seed(0))
    Vo = V.clone().requires_grad_(True); To = tex.clone().requires_grad_(True)
    Vg = V.clone().to(dev).requires_grad_(True); Tg = tex.clone().to(dev).requires_grad_(True)
    so, do_, co = ro.render_silhouettes(Vo,Fc), ro.render_depth(Vo,Fc), ro.render(Vo,Fc,To)
    sg, dg, cg = rg.render_silhouettes(Vg,Fc.to(dev)), rg.render_depth(Vg,Fc.to(dev)), rg.render(Vg,Fc.to(dev),Tg)
    print('cube aa',aa,'size',size,'sil eq',torch.equal(so,sg.cpu()),'depth rel',rel(dg,do_),'rgb rel',rel(cg,co))
    H,W = so.shape[-2:]
    wgt = torch.linspace(-1,1,W)[None,None,:]*torch.linspace(0.5,1.5,H)[None,:,None]
    lo = (so*wgt).sum()+(do_*0.01*wgt).sum()+(co*wgt[:,None]).sum()
    lg = (sg*wgt.to(dev)).sum()+(dg*0.01*wgt.to(dev)).sum()+(cg*wgt.to(dev)[:,None]).sum()
    lo.backward(); lg.backward()
    print('   grad V rel',rel(Vg.grad,Vo.grad),'grad T rel',rel(Tg.grad,To.grad), Vo.grad.abs().sum().item())
# --- derender3d bridge, small scene
verts, faces = wl.vkitti_scene_batch(2, seed=3, H=96, W=320, focal=180.)
Fb = faces[None].expand(2,-1,-1).contiguous()
ang = wl.vkitti_viewing_angle(96, 180.)
for rt,name in ((RenderType.Silhouette,'sil'),(RenderType.Depth,'depth'),(RenderType.Normal,'normal')):
    Vo = verts.clone().requires_grad_(True); Vg = verts.clone().to(dev).requires_grad_(True)
    io = O.derender_render(Vo, Fb, (96,320), ang, rt)
    R = Renderer(image_size=(96,320), viewing_angle=ang)
    ig = R(Vg, Fb.to(dev).int(), render_type=rt)
    g = torch.randn(io.shape, generator=torch.Generator().manual_seed(1))
    (io*g).sum().backward(); (ig*g.to(dev)).sum().backward()
    print(name, 'img rel', rel(ig,io), 'eq', torch.equal(ig.cpu(),io), 'grad rel', rel(Vg.grad, Vo.grad), 'cov', (io!=io[0,0,0,0]).float().mean().item())
# multi
Vg = verts.clone().to(dev).requires_grad_(True)
R = Renderer(image_size=(96,320), viewing_angle=ang)
out = R.forward_multi(Vg, Fb.to(dev).int())
g = torch.Generator().manual_seed(1)
gs=torch.randn(out['silhouette'].shape,generator=g); gd=torch.randn(out['depth'].shape,generator=g); gn=torch.randn(out['normal'].shape,generator=g)
((out['silhouette']*gs.to(dev)).sum()+(out['depth']*gd.to(dev)).sum()+(out['normal']*gn.to(dev)).sum()).backward()
Vo = verts.clone().requires_grad_(True)
lo = (O.derender_render(Vo,Fb,(96,320),ang,RenderType.Silhouette)*gs).sum()+(O.derender_render(Vo,Fb,(96,320),ang,RenderType.Depth)*gd).sum()+(O.derender_render(Vo,Fb,(96,320),ang,RenderType.Normal)*gn).sum()
lo.backward()
print('multi grad rel', rel(Vg.grad, Vo.grad))
# face index exactness
st = O.raster_forward
torch.cuda.synchronize()
# timing cfg2
verts, faces = wl.vkitti_scene_batch(32, seed=0)
Vg = verts.to(dev).requires_grad_(True); Fg = faces.to(dev).int()
R = Renderer(image_size=(375,1242), viewing_angle=wl.vkitti_viewing_angle())
gen = torch.Generator().manual_seed(1)
gs=torch.randn(32,1,375,1242,generator=gen).to(dev); gd=torch.randn(32,1,375,1242,generator=gen).to(dev); gn=torch.randn(32,3,375,1242,generator=gen).to(dev)
for fast in (False, True):
  for it in range(3):
    torch.cuda.synchronize(); t0=time.time()
    out = R.forward_multi(Vg, Fg, fast_math=fast)
    torch.cuda.synchronize(); t1=time.time()
    loss=(out['silhouette']*gs).sum()+(out['depth']*gd).sum()+(out['normal']*gn).sum()
    torch.cuda.synchronize(); t2=time.time()
    loss.backward()
    torch.cuda.synchronize(); t3=time.time()
    print('cfg2 fast',fast,'fwd ms',(t1-t0)*1e3,'bwd ms (incl torch loss bwd)',(t3-t2)*1e3, 'cov', (out['silhouette']>0).float().mean().item())
