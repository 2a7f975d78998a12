"""Per-parameter gradient error of the native training trunk against torchvision fp32; torch bf16 autocast as a yardstick."""
import copy, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '3d-sdn_b200'))
import torch
from derender3d.models.derenderer import Derenderer
from derender3d.models.encoder_tc import resnet18_trunk_train

def rel(a, b):
    return float((a.detach().float() - b.detach().float()).norm() / (b.detach().float().norm() + 1e-30))

dev = torch.device('cuda')
torch.manual_seed(0)
ref = Derenderer().to(dev).train()
mine, auto = copy.deepcopy(ref), copy.deepcopy(ref)
N = int(os.environ.get('N', 8))
x = torch.randn(N, 3, 224, 224, generator=torch.Generator().manual_seed(3)).to(dev)
t = torch.randn(N, 512, generator=torch.Generator().manual_seed(4)).to(dev)
f_ref = ref._trunk(x); (f_ref * t).sum().backward()
f_mine = resnet18_trunk_train(mine.net, x); (f_mine * t).sum().backward()
with torch.autocast('cuda', dtype=torch.bfloat16):
    f_auto = auto._trunk(x.contiguous(memory_format=torch.channels_last))
(f_auto.float() * t).sum().backward()
print('features: mine %.3e  autocast %.3e' % (rel(f_mine, f_ref), rel(f_auto, f_ref)))
for (n, p), (_, q), (_, a) in zip(mine.net.named_parameters(), ref.net.named_parameters(), auto.net.named_parameters()):
    if n.startswith('fc.'):
        continue
    print('%-32s mine %.3e  autocast %.3e   |g| %.3e' % (n, rel(p.grad, q.grad), rel(a.grad, q.grad), float(q.grad.norm())))
