"""GPU debug: last residual block's BN backward vs a CPU torch emulation."""
import os, sys
import torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dmc_oracle as O
from dmcnet_b200.engine import DmcEngine
from dmcnet_b200.trainer import FusedTrainStep, HParams

sd = O.build_state(51, None, seed=1)
flow, mv, res, target = O.make_inputs(2, 3, 51, seed=0)
st = {k: (v.clone().requires_grad_(True) if not O.is_buffer(k) else v.clone()) for k, v in sd.items()}
mvf = mv.reshape(-1, 2, 224, 224); resf = res.reshape(-1, 3, 224, 224)
gen = O.gen_tiny_forward(st, torch.cat((mvf, resf), 1)) + mvf
x = gen.detach(); p = 'base_model'
x = F.conv2d(x, st[p + '.conv1.weight'], None, 2, 3)
x = F.relu(O._bn(x, st, p + '.bn1', True, 1e-5)); x = F.max_pool2d(x, 3, 2, 1)
for li, (width, stride) in enumerate(O.RESNET18_STAGES, start=1):
    for b in range(2):
        q = '%s.layer%d.%d' % (p, li, b); s = stride if b == 0 else 1
        out = F.relu(O._bn(F.conv2d(x, st[q + '.conv1.weight'], None, s, 1), st, q + '.bn1', True, 1e-5))
        y2 = F.conv2d(out, st[q + '.conv2.weight'], None, 1, 1); y2.retain_grad()
        out = O._bn(y2, st, q + '.bn2', True, 1e-5)
        if (q + '.downsample.0.weight') in st:
            idt = O._bn(F.conv2d(x, st[q + '.downsample.0.weight'], None, s, 0), st, q + '.downsample.1', True, 1e-5)
        else:
            idt = x
        x = F.relu(out + idt); x.retain_grad()
xo, y2l = x, y2
pooled = F.adaptive_avg_pool2d(x, 1).flatten(1); pooled.retain_grad()
logits = F.linear(pooled, st[p + '.fc.weight'], st[p + '.fc.bias'])
loss = F.cross_entropy(logits.view(-1, 3, 51).mean(1), target)
loss.backward()

eng = DmcEngine(51, 3, 6)
eng.load_state(sd)
eng.debug_capture = True
tr = FusedTrainStep(eng, HParams(), 2)
tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda(), apply=False)
torch.cuda.synchronize()
d = {k: v.cpu() for k, v in eng.debug.items()}

def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))

def to_nchw(t, C):   # [P][C] padded pixel-major -> [N,C,7,7]
    return t.view(6, 9, 9, C)[:, 1:8, 1:8, :].permute(0, 3, 1, 2)

print('d_pooled', rel(d['d_pooled'], pooled.grad))
print('g_a(dOut)', rel(to_nchw(d['g_a'], 512), xo.grad))
print('act', rel(to_nchw(d['act'], 512), xo.detach()))
print('Y2', rel(to_nchw(d['Y'], 512), y2l.detach()))
dz_ref = xo.grad * (xo > 0)
print('dz', rel(to_nchw(d['dz'], 512), dz_ref))
print('sum dz', rel(d['sums2'][0].float(), dz_ref.sum((0, 2, 3))))
print('G(dY2)', rel(to_nchw(d['G'], 512), y2l.grad))
mean = y2l.detach().mean((0, 2, 3)); var = y2l.detach().var((0, 2, 3), unbiased=False)
print('mean', rel(d['mean'], mean), 'invstd', rel(d['invstd'], 1 / torch.sqrt(var + 1e-5)))
bad = (to_nchw(d['dz'], 512) - dz_ref).abs()
print('dz worst idx', torch.nonzero(bad == bad.max())[:4].tolist(), bad.max().item(), dz_ref.abs().max().item())
border = d['dz'].view(6, 9, 9, 512).clone(); border[:, 1:8, 1:8, :] = 0
print('dz border max', border.abs().max().item())
