"""Developer diagnostic (GPU): per-layer forward error of the ContextNetwork plan against the oracle's
functions evaluated in float64, and the gen_flow error that feeds the classifier."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dmc_oracle as O                      # noqa: E402  (checker)
from dmcnet_b200.engine import DmcEngine, CONTEXT_RING   # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    n = 3 * B
    sd = O.build_state(51, None, seed=1, arch_estimator='ContextNetwork')
    flow, mv, res, target = O.make_inputs(B, 3, 51, seed=0)
    eng = DmcEngine(51, 3, n, arch_estimator='ContextNetwork')
    eng.load_state(sd)
    eng.forward(mv.cuda(), res.cuda(), train=True)
    x = torch.cat((mv.reshape(n, 2, 224, 224), res.reshape(n, 3, 224, 224)), 1).double()
    R = CONTEXT_RING
    for dt in (torch.float64, torch.float32):
        h = x.to(dt)
        for i, (co, d) in enumerate(O.context_layers(0, 0)):
            p = 'gen_flow_model.conv_context.%d' % i
            y = F.conv2d(h, sd[p + '.0.weight'].to(dt), None, 1, d, d)
            h = F.leaky_relu(F.batch_norm(y, None, None, sd[p + '.1.weight'].to(dt), sd[p + '.1.bias'].to(dt), True, 0.1, 1e-5), 0.1)
            L = eng.ctx_layers[i]
            Hp, Wp = eng.ctx_geo
            Y = L['Y'].view(n, Hp, Wp, L['Np'])[:, R:, R:, :co].permute(0, 3, 1, 2).double().cpu()
            A = (L['act_hi'].float() + L['act_lo'].float()).view(n, Hp, Wp, L['Np'])[:, R:, R:, :co].permute(0, 3, 1, 2).double().cpu()
            ey = float((Y - y.double()).abs().max() / y.double().abs().max())
            ea = float((A - h.double()).abs().max() / h.double().abs().max())
            print('%s layer %d (cout %3d dil %2d): conv out rel err %.2e   act rel err %.2e' % (str(dt)[6:], i, co, d, ey, ea))
        gf = h.double() + mv.reshape(n, 2, 224, 224).double()
        print('%s gen_flow rel err %.2e' % (str(dt)[6:], float((eng.gen_flow.double().cpu() - gf).abs().max() / gf.abs().max())))


if __name__ == '__main__':
    main()
