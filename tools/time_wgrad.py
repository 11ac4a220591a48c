"""Times the tensor-core weight-gradient GEMM on the ResNet-18 layer shapes of the bench (192 frames)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcnet_b200 import ops
from dmcnet_b200.engine import _taps_s1

N = 192
def main():
    for c, h in ((64, 56), (128, 28), (256, 14), (512, 7)):
        Hp = ops.padded(h)
        P = N * Hp * Hp
        G = torch.randn(P, c, device='cuda')
        X = torch.randn(P, c, device='cuda')
        Gh = G.to(torch.bfloat16); Gl = (G - Gh.float()).to(torch.bfloat16)
        Xh = X.to(torch.bfloat16); Xl = (X - Xh.float()).to(torch.bfloat16)
        dW = torch.zeros(c, c, 3, 3, device='cuda')
        shift, phase, bsel = _taps_s1(Hp)
        ws = torch.empty(ops.wgrad_workspace_floats(P, c, c, 9), device='cuda') if os.environ.get('WS') else None
        fn = lambda: ops.wgrad_gemm(Gh, Gl, Xh[None], Xl[None], dW, P=P, Cout=c, x_phases=1, Cin=c, shift=shift,
                                    phase=phase, bsel=bsel, engine='tc', oihw_taps=9, workspace=ws)
        fn(); torch.cuda.synchronize()
        ref = dW.clone(); dW.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 10 * 1e-3
        fl = 2.0 * N * h * h * c * c * 9
        # fp64 check of one tap on a sample
        chk = (G.double().t() @ X.double())          # centre tap
        err = float((ref[:, :, 1, 1].double() - chk).abs().max() / chk.abs().max())
        print('C=%3d %2dx%2d  %7.1f us  %6.1f TFLOP/s  centre-tap rel err %.2e' % (c, h, h, t * 1e6, fl / t / 1e12, err), flush=True)

if __name__ == '__main__':
    main()
