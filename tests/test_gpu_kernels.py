"""GPU: every kernel family against a plain PyTorch fp32 (CPU) reference of the
same op on identical inputs, through the C-ABI.  Tolerances: 2e-5 for fp32
CUDA-core kernels, 5e-5 for the bf16x3 tensor-core GEMMs (~2^-16 per product),
relative to the largest reference magnitude."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from dmcnet_b200 import ops


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30))


def split(x):
    hi = x.to(torch.bfloat16)
    return hi.contiguous(), (x - hi.float()).to(torch.bfloat16).contiguous()


def to_pixel(x):      # [N,C,H,W] -> padded pixel-major [N*(H+2)*(W+2), C]
    n, c, h, w = x.shape
    p = torch.zeros(n, ops.padded(h), ops.padded(w), c)
    p[:, 1:h + 1, 1:w + 1, :] = x.permute(0, 2, 3, 1)
    return p.reshape(-1, c).contiguous()


def from_pixel(flat, n, c, h, w):
    return flat.reshape(n, ops.padded(h), ops.padded(w), c)[:, 1:h + 1, 1:w + 1, :].permute(0, 3, 1, 2)


# ------------------------------------------------------------------ planar small-channel convs
CONV_CASES = [(2, 33, 2, 64, 64, 3, 1), (2, 5, 8, 64, 96, 3, 1), (1, 21, 6, 40, 36, 3, 1),
              (2, 27, 4, 32, 64, 3, 1), (2, 16, 16, 56, 56, 3, 1), (2, 2, 16, 64, 64, 3, 2),
              (2, 32, 64, 28, 28, 3, 2), (2, 2, 64, 64, 64, 7, 2), (1, 3, 8, 17, 20, 3, 1)]


@pytest.mark.parametrize('N,Cin,Cout,H,W,ks,stride', CONV_CASES)
def test_planar_conv_fwd_dgrad_wgrad(N, Cin, Cout, H, W, ks, stride):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Cin, H, W, generator=g, requires_grad=True)
    w = (torch.randn(Cout, Cin, ks, ks, generator=g) * 0.2).requires_grad_(True)
    b = torch.randn(Cout, generator=g, requires_grad=True)
    mask = torch.empty(N, Cout).bernoulli_(0.75, generator=g) / 0.75
    pad = ks // 2
    pre = F.conv2d(x, w, b, stride, pad)
    y = F.leaky_relu(pre, 0.2) * mask.view(N, Cout, 1, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    Ho, Wo = y.shape[-2:]
    xd, wd, bd = x.detach().cuda(), w.detach().cuda(), b.detach().cuda()
    yo = torch.full((N, Cout, Ho, Wo), float('nan'), device='cuda')
    ops.conv_fwd(xd, Cin * H * W, Cin, H, W, wd, bd, Cout, ks, stride, yo, Cout * Ho * Wo, N, slope=0.2,
                 mask=mask.cuda())
    assert rel(yo, y.detach()) < 2e-5
    dpre = torch.empty_like(yo)
    ops.act_bwd_planar(dy.cuda(), Cout * Ho * Wo, yo, Cout * Ho * Wo, mask.cuda(), 0.2, Cout, Ho * Wo, N,
                       dpre, Cout * Ho * Wo)
    dW, dB = torch.zeros_like(wd), torch.zeros_like(bd)
    ops.conv_wgrad(xd, Cin * H * W, Cin, H, W, dpre, Cout * Ho * Wo, Cout, ks, stride, dW, dB, N)
    assert rel(dW, w.grad) < 2e-5 and rel(dB, b.grad) < 2e-5
    dX = torch.full((N, Cin, H, W), float('nan'), device='cuda')
    ops.conv_dgrad(dpre, Cout * Ho * Wo, Cout, wd, Cin, Cin, ks, stride, dX, Cin * H * W, H, W, N)
    assert rel(dX, x.grad) < 2e-5
    if ks == 3 and stride == 1:       # the engine's stride-1 path: forward conv with flipped weights
        wT = torch.zeros(Cin * Cout * 9, device='cuda')
        ops.weight_flip(wd, Cout, Cin, Cin, wT)
        dX2 = torch.full((N, Cin, H, W), float('nan'), device='cuda')
        ops.conv_fwd(dpre, Cout * Ho * Wo, Cout, H, W, wT, None, Cin, 3, 1, dX2, Cin * H * W, N)
        assert rel(dX2, x.grad) < 2e-5


def test_conv_fwd_channel_subrange_add_and_accumulate():
    """Reads / writes channel sub-ranges of a wider buffer (the dense concat), '+ input_mv', and '+='."""
    g = torch.Generator().manual_seed(1)
    N, H, W = 2, 32, 32
    buf = torch.randn(N, 10, H, W, generator=g)
    w = torch.randn(3, 6, 3, 3, generator=g) * 0.3
    b = torch.randn(3, generator=g)
    add = torch.randn(N, 3, H, W, generator=g)
    ref = buf.clone()
    ref[:, 1:4] = ref[:, 1:4] + F.conv2d(buf[:, 4:10], w, b, 1, 1) + add
    d = buf.cuda()
    flat = d.view(-1)
    ops.conv_fwd(flat[4 * H * W:], 10 * H * W, 6, H, W, w.cuda(), b.cuda(), 3, 3, 1, flat[1 * H * W:],
                 10 * H * W, N, slope=1.0, add=add.cuda(), add_ns=3 * H * W, accumulate=True)
    assert rel(d, ref) < 2e-5


def test_stride2_backward_through_space_to_depth():
    """s2d / d2s / weight maps (the discriminator's stride-2 backward path) vs torch autograd of the
    stride-2 conv itself."""
    g = torch.Generator().manual_seed(7)
    N, ci, co, H, W = 3, 4, 8, 32, 48
    x = torch.randn(N, ci, H, W, generator=g, requires_grad=True)
    w = (torch.randn(co, ci, 3, 3, generator=g) * 0.2).requires_grad_(True)
    y = F.conv2d(x, w, None, 2, 1)
    dy = torch.randn(y.shape, generator=g)
    gx, gw = torch.autograd.grad(y, (x, w), dy)
    Ho, Wo, c4 = H // 2, W // 2, 4 * ci
    xd, wd, dyd = x.detach().cuda(), w.detach().cuda(), dy.cuda()
    S = torch.full((N, c4, Ho, Wo), float('nan'), device='cuda')
    ops.s2d_planar(xd, ci * H * W, ci, H, W, S, ci * H * W, N)
    ref_S = torch.cat([x.detach()[:, :, pr::2, pc::2] for pr in (0, 1) for pc in (0, 1)], 1)
    assert torch.equal(S.cpu(), ref_S)
    # weight gradient: stride-1 wgrad on S, gathered back into OIHW (accumulating)
    dW3 = torch.zeros(co, c4, 3, 3, device='cuda')
    ops.conv_wgrad(S, ci * H * W, c4, Ho, Wo, dyd, co * Ho * Wo, co, 3, 1, dW3, None, N)
    dW = torch.ones(co, ci, 3, 3, device='cuda')
    ops.s2_weight_map(dW3, dW, co, ci, False)
    assert rel(dW - 1.0, gw) < 2e-5
    # data gradient: conv of dY with the flipped space-to-depth weight, then depth-to-space
    W3 = torch.full((co, c4, 3, 3), float('nan'), device='cuda')
    ops.s2_weight_map(wd, W3, co, ci, True)
    W3t = torch.zeros(c4 * co * 9, device='cuda')
    ops.weight_flip(W3, co, c4, c4, W3t)
    # forward through the 2x2-tap kernel (taps {0,1}), with bias / LeakyReLU / channel mask epilogue
    b = torch.randn(co, generator=g)
    mask = torch.empty(N, co).bernoulli_(0.75, generator=g) / 0.75
    ref_y = F.leaky_relu(y.detach() + b.view(1, co, 1, 1), 0.2) * mask.view(N, co, 1, 1)
    Yo = torch.full((N, co, Ho, Wo), float('nan'), device='cuda')
    ops.conv3x3_taps2(S, ci * H * W, c4, Ho, Wo, W3, b.cuda(), co, 0, Yo, co * Ho * Wo, N, slope=0.2,
                      mask=mask.cuda())
    assert rel(Yo, ref_y) < 2e-5
    dS = torch.full((N, c4, Ho, Wo), float('nan'), device='cuda')
    ops.conv3x3_taps2(dyd, co * Ho * Wo, co, Ho, Wo, W3t, None, c4, 1, dS, ci * H * W, N)
    dX = torch.ones(N, ci, H, W, device='cuda')
    ops.d2s_planar(dS, ci * H * W, ci, H, W, dX, ci * H * W, N, accumulate=True)
    assert rel(dX - 1.0, gx) < 2e-5
    ops.d2s_planar(dS, ci * H * W, ci, H, W, dX, ci * H * W, N)
    assert rel(dX, gx) < 2e-5


# ------------------------------------------------------------------ tensor-core tap GEMMs
@pytest.mark.parametrize('engine', ['tc', 'simt'])
@pytest.mark.parametrize('n,cin,cout,h,stride', [(3, 64, 64, 12, 1), (2, 128, 256, 8, 1), (2, 64, 128, 16, 2),
                                                  (5, 256, 512, 14, 2)])
def test_tap_gemm_conv_fprop_dgrad_wgrad(engine, n, cin, cout, h, stride):
    from dmcnet_b200.engine import _taps_s1, _taps_s2
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n, cin, h, h, generator=g, requires_grad=True)
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05).requires_grad_(True)
    y = F.conv2d(x, w, None, stride, 1)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    ho = h // stride
    Hp = Wp = ops.padded(ho)
    P = n * Hp * Wp
    W_hi = torch.zeros(9, cout, cin, dtype=torch.bfloat16, device='cuda')
    W_lo, Wt_hi, Wt_lo = (torch.zeros_like(W_hi) for _ in range(3))
    ops.weight_prep(w.detach().cuda().contiguous(), cout, cin, 9, W_hi, W_lo, Wt_hi.view(9, cin, cout),
                    Wt_lo.view(9, cin, cout))
    if stride == 1:
        A = to_pixel(x.detach())[None]
        shift, phase, bsel = _taps_s1(Wp)
        phases = 1
    else:
        hi_, lo_ = split(to_pixel(x.detach()).cuda())
        xp_hi = torch.zeros(4, P, cin, dtype=torch.bfloat16, device='cuda')
        xp_lo = torch.zeros_like(xp_hi)
        ops.phase_split(hi_, lo_, n, h, h, cin, xp_hi, xp_lo)
        shift, phase, bsel = _taps_s2(Wp)
        phases = 4
    if stride == 1:
        A_hi, A_lo = split(A.cuda())
    else:
        A_hi, A_lo = xp_hi, xp_lo
    D = torch.full((P, cout), float('nan'), device='cuda')
    stats = torch.zeros(2, cout, dtype=torch.float64, device='cuda')
    ops.tap_gemm(A_hi, A_lo, W_hi, W_lo, D, a_phases=phases, a_rows=P, K=cin, b_slices=9, N=cout, M=P,
                 ldD=cout, Hp=Hp, Wp=Wp, shift=shift, phase=phase, bsel=bsel, engine=engine, stats=stats)
    assert rel(from_pixel(D, n, cout, ho, ho), y.detach()) < 5e-5
    # fused BatchNorm batch statistics (per-channel sum and sum of squares of the conv output)
    yd = y.detach().double()
    assert rel(stats[0], yd.sum((0, 2, 3))) < 5e-5 * (n * ho * ho) ** 0.5
    assert rel(stats[1], (yd * yd).sum((0, 2, 3))) < 5e-5
    ring = D.view(n, Hp, Wp, cout).clone()
    ring[:, 1:ho + 1, 1:ho + 1] = 0
    assert float(ring.abs().max()) == 0.0                         # epilogue keeps the zero ring
    # wgrad
    G_hi, G_lo = split(to_pixel(dy).cuda())
    dWs = torch.zeros(9, cout, cin, device='cuda')
    ops.wgrad_gemm(G_hi, G_lo, A_hi, A_lo, dWs, P=P, Cout=cout, x_phases=phases, Cin=cin, shift=shift,
                   phase=phase, bsel=bsel, engine=engine)
    gw = torch.empty(cout, cin, 3, 3, device='cuda')
    ops.wgrad_unpack(dWs, gw, cout, cin, 9)
    assert rel(gw, w.grad) < 5e-5
    if engine == 'tc':
        # workspace path: partial tiles + fixed-order reduction straight into OIHW; bit-reproducible
        ws = torch.empty(ops.wgrad_workspace_floats(P, cout, cin, 9), device='cuda')
        runs = []
        for _ in range(2):
            gw2 = torch.zeros(cout, cin, 3, 3, device='cuda')
            ops.wgrad_gemm(G_hi, G_lo, A_hi, A_lo, gw2, P=P, Cout=cout, x_phases=phases, Cin=cin, shift=shift,
                           phase=phase, bsel=bsel, engine=engine, oihw_taps=9, workspace=ws)
            runs.append(gw2)
        assert rel(runs[0], w.grad) < 5e-5
        assert torch.equal(runs[0], runs[1])
    # dgrad
    if stride == 1:
        dX = torch.full((P, cin), float('nan'), device='cuda')
        ops.tap_gemm(G_hi, G_lo, Wt_hi, Wt_lo, dX, a_phases=1, a_rows=P, K=cout, b_slices=9, N=cin, M=P,
                     ldD=cin, Hp=Hp, Wp=Wp, shift=[-s for s in shift], phase=phase, bsel=bsel, engine=engine)
        assert rel(from_pixel(dX, n, cin, h, h), x.grad) < 5e-5
        # lo plane of the gradient omitted (NULL): the kernels must then equal the exact result
        # for dY rounded to bf16 -- the engine's opt-in grad_bf16 mode
        dy_r = dy.to(torch.bfloat16).float()
        y2 = F.conv2d(x, w, None, stride, 1)
        gx_r, gw_r = torch.autograd.grad(y2, (x, w), dy_r)
        dX.fill_(float('nan'))
        ops.tap_gemm(G_hi, None, Wt_hi, Wt_lo, dX, a_phases=1, a_rows=P, K=cout, b_slices=9, N=cin, M=P,
                     ldD=cin, Hp=Hp, Wp=Wp, shift=[-s for s in shift], phase=phase, bsel=bsel, engine=engine)
        assert rel(from_pixel(dX, n, cin, h, h), gx_r) < 5e-5
        dWs.zero_()
        ops.wgrad_gemm(G_hi, None, A_hi, A_lo, dWs, P=P, Cout=cout, x_phases=phases, Cin=cin, shift=shift,
                       phase=phase, bsel=bsel, engine=engine)
        ops.wgrad_unpack(dWs, gw, cout, cin, 9)
        assert rel(gw, gw_r) < 5e-5
    else:
        dxp = torch.zeros(4, P, cin, device='cuda')
        for ph in range(4):
            sh = [-shift[t] for t in range(9) if phase[t] == ph]
            bs = [bsel[t] for t in range(9) if phase[t] == ph]
            ops.tap_gemm(G_hi, G_lo, Wt_hi, Wt_lo, dxp[ph], a_phases=1, a_rows=P, K=cout, b_slices=9, N=cin,
                         M=P, ldD=cin, Hp=Hp, Wp=Wp, shift=sh, phase=[0] * len(sh), bsel=bs, engine=engine)
        dX = torch.empty(n * ops.padded(h) * ops.padded(h), cin, device='cuda')
        ops.phase_unsplit(dxp, n, h, h, cin, dX)
        assert rel(from_pixel(dX, n, cin, h, h), x.grad) < 5e-5


# ------------------------------------------------------------------ BatchNorm (pixel-major and planar)
def test_pixel_bn_train_forward_backward_with_residual():
    g = torch.Generator().manual_seed(3)
    n, c, h = 4, 64, 10
    y = (torch.randn(n, c, h, h, generator=g) * 2 + 0.5).requires_grad_(True)
    res = torch.randn(n, c, h, h, generator=g).relu()
    gamma = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    beta = torch.randn(c, generator=g, requires_grad=True)
    rm, rv = torch.zeros(c), torch.ones(c)
    out = F.relu(F.batch_norm(y, rm, rv, gamma, beta, True, 0.1, 1e-5) + res)
    dout = torch.randn(out.shape, generator=g)
    out.backward(dout)
    Hp = ops.padded(h)
    P = n * Hp * Hp
    dev = dict(device='cuda')
    Y = to_pixel(y.detach()).cuda()
    sums = torch.zeros(2, c, dtype=torch.float64, **dev)
    scale, shift, mean, invstd = (torch.zeros(c, **dev) for _ in range(4))
    rmd, rvd = torch.zeros(c, **dev), torch.ones(c, **dev)
    nbt = torch.zeros((), dtype=torch.int64, **dev)
    ops.bn_stats(Y, P, c, sums)
    ops.bn_finalize(sums, float(n * h * h), gamma.detach().cuda(), beta.detach().cuda(), rmd, rvd, nbt, 0.1,
                    1e-5, c, scale, shift, mean, invstd)
    assert rel(rmd, rm) < 1e-5 and rel(rvd, rv) < 1e-5 and int(nbt) == 1
    r_hi, r_lo = split(to_pixel(res).cuda())
    o_hi = torch.zeros(P, c, dtype=torch.bfloat16, **dev)
    o_lo = torch.zeros_like(o_hi)
    ops.bn_apply(Y, scale, shift, P, c, Hp, Hp, True, o_hi, o_lo, res_hi=r_hi, res_lo=r_lo)
    got = from_pixel(o_hi.float() + o_lo.float(), n, c, h, h)
    assert rel(got, out.detach()) < 2e-5
    # backward
    gA = to_pixel(dout).cuda()
    sums2 = torch.zeros(2, c, dtype=torch.float64, **dev)
    G_hi = torch.zeros(P, c, dtype=torch.bfloat16, **dev)
    G_lo = torch.zeros_like(G_hi)
    dz = torch.zeros(P, c, **dev)
    dgam, dbet = torch.zeros(c, **dev), torch.zeros(c, **dev)
    ops.bn_bwd_reduce(gA, None, o_hi, Y, mean, invstd, P, c, Hp, Hp, sums2)
    ops.bn_bwd_apply(gA, None, o_hi, Y, mean, invstd, gamma.detach().cuda(), sums2, float(n * h * h), P, c,
                     Hp, Hp, G_hi, G_lo, dz, dgam, dbet)
    assert rel(from_pixel(G_hi.float() + G_lo.float(), n, c, h, h), y.grad) < 2e-5
    assert rel(dgam, gamma.grad) < 2e-5 and rel(dbet, beta.grad) < 2e-5
    assert rel(from_pixel(dz, n, c, h, h), dout * (out.detach() > 0)) < 1e-6


def test_planar_bn_eps08_forward_backward():
    """Discriminator BatchNorm2d(C, 0.8): the 0.8 is eps (code/dmcnet_GAN/model.py:264)."""
    g = torch.Generator().manual_seed(4)
    n, c, h = 6, 16, 14
    x = torch.randn(n, c, h, h, generator=g, requires_grad=True)
    gamma = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    beta = torch.randn(c, generator=g, requires_grad=True)
    out = F.batch_norm(x, torch.zeros(c), torch.ones(c), gamma, beta, True, 0.1, 0.8)
    dout = torch.randn(out.shape, generator=g)
    out.backward(dout)
    dev = dict(device='cuda')
    X = x.detach().cuda()
    sums = torch.zeros(2, c, dtype=torch.float64, **dev)
    scale, shift, mean, invstd = (torch.zeros(c, **dev) for _ in range(4))
    ops.bn_stats_planar(X, c * h * h, c, h * h, n, sums)
    ops.bn_finalize(sums, float(n * h * h), gamma.detach().cuda(), beta.detach().cuda(), None, None, None,
                    0.1, 0.8, c, scale, shift, mean, invstd)
    Z = torch.empty_like(X)
    ops.bn_apply_planar(X, c * h * h, scale, shift, c, h * h, n, False, Z, c * h * h)
    assert rel(Z, out.detach()) < 2e-5
    sums2 = torch.zeros(2, c, dtype=torch.float64, **dev)
    dX = torch.empty_like(X)
    dgam, dbet = torch.zeros(c, **dev), torch.zeros(c, **dev)
    ops.bn_bwd_reduce_planar(dout.cuda(), c * h * h, X, c * h * h, mean, invstd, c, h * h, n, sums2)
    ops.bn_bwd_apply_planar(dout.cuda(), c * h * h, X, c * h * h, mean, invstd, gamma.detach().cuda(), sums2,
                            float(n * h * h), c, h * h, n, dX, c * h * h, dgam, dbet)
    assert rel(dX, x.grad) < 2e-5 and rel(dgam, gamma.grad) < 2e-5 and rel(dbet, beta.grad) < 2e-5


@pytest.mark.parametrize('N,H,W', [(2, 64, 64), (3, 224, 224), (1, 40, 48)])
def test_stem_conv_tensor_core_forward_and_wgrad(N, H, W):
    """7x7/2 stem conv through the im2col-in-shared-memory tcgen05 kernel vs torch fp32."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, 2, H, W, generator=g)
    w = torch.randn(64, 2, 7, 7, generator=g) * 0.1
    y = F.conv2d(x, w, None, 2, 3)
    Y = torch.full((N, 64, H // 2, W // 2), float('nan'), device='cuda')
    wb = torch.zeros(128 * 128, dtype=torch.bfloat16, device='cuda')
    ops.stem_conv_tc_fwd(x.cuda(), 2 * H * W, H, W, w.cuda(), wb, Y, 64 * (H // 2) * (W // 2), N)
    assert rel(Y, y) < 2e-5
    # weight gradient through the same im2col tiles (deterministic: workspace + fixed-order reduce)
    wr = w.clone().requires_grad_(True)
    dy = torch.randn(y.shape, generator=g)
    F.conv2d(x, wr, None, 2, 3).backward(dy)
    ws = torch.empty(ops.stem_wgrad_workspace_floats(), device='cuda')
    runs = []
    for _ in range(2):
        dW = torch.zeros(64, 2, 7, 7, device='cuda')
        ops.stem_conv_tc_wgrad(x.cuda(), 2 * H * W, H, W, dy.cuda(), 64 * (H // 2) * (W // 2), dW, ws, N)
        runs.append(dW)
    assert rel(runs[0], wr.grad) < 2e-5
    assert torch.equal(runs[0], runs[1])


def test_stem_bn_relu_maxpool_forward_backward():
    g = torch.Generator().manual_seed(5)
    n, c, h = 3, 64, 32
    y = torch.randn(n, c, h, h, generator=g, requires_grad=True)
    scale = torch.rand(c, generator=g) + 0.5
    shift = torch.randn(c, generator=g) * 0.3
    a = F.relu(y * scale.view(1, c, 1, 1) + shift.view(1, c, 1, 1))
    p = F.max_pool2d(a, 3, 2, 1)
    dp = torch.randn(p.shape, generator=g)
    (da,) = torch.autograd.grad(p, a, dp, retain_graph=True)
    dz_ref = da * (a.detach() > 0)
    hq = h // 2
    dev = dict(device='cuda')
    o_hi = torch.zeros(n * ops.padded(hq) * ops.padded(hq), c, dtype=torch.bfloat16, **dev)
    o_lo = torch.zeros_like(o_hi)
    idx = torch.zeros(n * hq * hq * c, dtype=torch.uint8, **dev)
    Y = y.detach().cuda()
    ops.stem_pool_fwd(Y, scale.cuda(), shift.cuda(), n, c, h, h, o_hi, o_lo, idx)
    assert rel(from_pixel(o_hi.float() + o_lo.float(), n, c, hq, hq), p.detach()) < 2e-5
    dZ = torch.empty_like(Y)
    ops.stem_pool_bwd(to_pixel(dp).cuda(), None, idx, Y, scale.cuda(), shift.cuda(), n, c, h, h, dZ)
    assert rel(dZ, dz_ref) < 1e-6


# ------------------------------------------------------------------ heads and optimizer
def test_ce_head_consensus_loss_grad_topk():
    g = torch.Generator().manual_seed(6)
    B, S, C = 16, 3, 51
    logits = (torch.randn(B * S, C, generator=g) * 2).requires_grad_(True)
    target = torch.randint(0, C, (B,), generator=g)
    out = logits.view(B, S, C).mean(1)
    loss = F.cross_entropy(out, target)
    (loss * 0.7).backward()
    dev = dict(device='cuda')
    cons, dl, st = torch.zeros(B, C, **dev), torch.zeros(B * S, C, **dev), torch.zeros(4, **dev)
    ops.ce_head(logits.detach().cuda(), B, S, C, target.cuda(), 0.7 / B, cons, dl, st)
    st = st.cpu()
    assert rel(cons, out.detach()) < 1e-6 and rel(dl, logits.grad) < 1e-5
    assert abs(float(st[0]) / B - float(loss.detach())) < 1e-5
    _, pred = out.topk(5, 1, True, True)
    correct = pred.eq(target.view(-1, 1))
    assert int(st[1]) == int(correct[:, :1].sum()) and int(st[2]) == int(correct.sum())


def test_mse_head_and_linear():
    g = torch.Generator().manual_seed(7)
    a = torch.randn(6, 2, 32, 32, generator=g, requires_grad=True)
    b = torch.randn(6, 2, 32, 32, generator=g)
    loss = F.mse_loss(a, b)
    (loss * 10).backward()
    dev = dict(device='cuda')
    dgen, s = torch.zeros_like(a, **dev), torch.zeros(1, dtype=torch.float64, **dev)
    ops.mse_head(a.detach().cuda(), b.cuda(), a.numel(), 2.0 * 10 / a.numel(), dgen, s)
    assert abs(float(s.cpu()[0]) / a.numel() - float(loss)) < 1e-6 and rel(dgen, a.grad) < 1e-6
    x = torch.randn(12, 512, generator=g, requires_grad=True)
    w = (torch.randn(51, 512, generator=g) * 0.05).requires_grad_(True)
    bias = torch.randn(51, generator=g, requires_grad=True)
    y = F.linear(x, w, bias)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    yo = torch.zeros(12, 51, **dev)
    ops.linear_fwd(x.detach().cuda(), w.detach().cuda(), bias.detach().cuda(), 12, 512, 51, yo)
    dx, dw, db = torch.zeros(12, 512, **dev), torch.zeros(51, 512, **dev), torch.zeros(51, **dev)
    ops.linear_bwd(dy.cuda(), x.detach().cuda(), w.detach().cuda(), 12, 512, 51, dx, dw, db)
    assert rel(yo, y.detach()) < 1e-5 and rel(dx, x.grad) < 1e-5 and rel(dw, w.grad) < 1e-5
    assert rel(db, bias.grad) < 1e-5


def test_fused_adam_matches_torch_adam():
    """torch.optim.Adam(eps=1e-3) with one param group per tensor (code/dmcnet/train.py:121-142)."""
    g = torch.Generator().manual_seed(8)
    shapes = [(64, 2, 7, 7), (64,), (51, 512), (5,), (1030,)]
    lrs = [1e-4, 1e-4, 1e-2, 1e-2, 3e-3]
    wds = [1e-4, 0.0, 1e-4, 0.0, 1e-4]
    ps = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    opt = torch.optim.Adam([{'params': [p], 'lr': lr, 'weight_decay': wd} for p, lr, wd in zip(ps, lrs, wds)],
                           eps=1e-3)
    offs, off = [], 0
    for s in shapes:
        offs.append(off)
        off += (torch.Size(s).numel() + 63) // 64 * 64
    dev = dict(device='cuda')
    flat, gr, m, v = (torch.zeros(off, **dev) for _ in range(4))
    chunks = []
    for ti, (s, o) in enumerate(zip(shapes, offs)):
        n = torch.Size(s).numel()
        flat[o:o + n] = ps[ti].detach().reshape(-1).cuda()
        for c0 in range(0, n, 1024):
            chunks.append((o + c0, min(1024, n - c0), ti, 0))
    ch = torch.tensor(chunks, dtype=torch.int32).cuda()
    hyper = torch.tensor(list(zip(lrs, wds)), dtype=torch.float32).cuda()
    step = torch.zeros(1, dtype=torch.int32, **dev)
    for it in range(3):
        for ti, p in enumerate(ps):
            p.grad = torch.randn(p.shape, generator=g) * (10.0 ** (-it))
            n = p.numel()
            gr[offs[ti]:offs[ti] + n] = p.grad.reshape(-1).cuda()
        opt.step()
        ops.adam_step(flat, gr, m, v, ch, len(chunks), hyper.view(-1), step, 0.9, 0.999, 1e-3)
        for ti, p in enumerate(ps):
            n = p.numel()
            assert rel(flat[offs[ti]:offs[ti] + n], p.detach().reshape(-1)) < 1e-6, (it, ti)
    assert int(step) == 3


# ------------------------------------------------------------------ full-size tensor-core GEMMs
# The layer shapes of BASELINE config 2 (192 frames): thousands of M tiles per launch, so the
# persistent loop of tap_gemm_ws_kernel (TMEM accumulator ping-pong over ~33 tiles per CTA, the
# statistics flush at every change of the output-channel block) and the multi-split wgrad plans
# run exactly as in the bench.  Reference: torch fp64 on the same device (checker only).
def _pixel_cuda(x):      # [N,C,H,W] cuda -> padded pixel-major [N*Hp*Wp, C] fp32 cuda
    n, c, h, w = x.shape
    p = torch.zeros(n, ops.padded(h), ops.padded(w), c, device=x.device)
    p[:, 1:h + 1, 1:w + 1, :] = x.permute(0, 2, 3, 1).float()
    return p.reshape(-1, c).contiguous()


def _from_pixel_cuda(flat, n, c, h, w):
    return flat.reshape(n, ops.padded(h), ops.padded(w), c)[:, 1:h + 1, 1:w + 1, :].permute(0, 3, 1, 2)


FULL_CASES = [(192, 64, 64, 56, 1), (192, 128, 128, 28, 1), (192, 256, 512, 14, 2), (192, 512, 512, 7, 1),
              (192, 64, 128, 56, 2)]


@pytest.mark.parametrize('n,cin,cout,h,stride', FULL_CASES)
def test_tap_gemm_full_size_layers_vs_fp64(n, cin, cout, h, stride):
    from dmcnet_b200.engine import _taps_s1, _taps_s2
    dev = torch.device('cuda')
    g = torch.Generator(device='cuda').manual_seed(11)
    x = torch.randn(n, cin, h, h, generator=g, device=dev)
    w = torch.randn(cout, cin, 3, 3, generator=g, device=dev) * 0.05
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    y = F.conv2d(xd, wd, None, stride, 1)
    dy = torch.randn(y.shape, generator=g, device=dev)
    y.backward(dy.double())
    ho = h // stride
    Hp = Wp = ops.padded(ho)
    P = n * Hp * Wp
    assert (P // 128) * (cout // 128 if cout >= 128 else 1) > 148     # more tiles than CTAs: persistent loop
    W_hi = torch.zeros(9, cout, cin, dtype=torch.bfloat16, device=dev)
    W_lo, Wt_hi, Wt_lo = (torch.zeros_like(W_hi) for _ in range(3))
    ops.weight_prep(w.contiguous(), cout, cin, 9, W_hi, W_lo, Wt_hi.view(9, cin, cout), Wt_lo.view(9, cin, cout))
    hi_, lo_ = split(_pixel_cuda(x))
    if stride == 1:
        A_hi, A_lo, phases = hi_, lo_, 1
        shift, phase, bsel = _taps_s1(Wp)
    else:
        A_hi = torch.zeros(4, P, cin, dtype=torch.bfloat16, device=dev)
        A_lo = torch.zeros_like(A_hi)
        ops.phase_split(hi_, lo_, n, h, h, cin, A_hi, A_lo)
        shift, phase, bsel = _taps_s2(Wp)
        phases = 4
    D = torch.full((P, cout), float('nan'), device=dev)
    stats = torch.zeros(2, cout, dtype=torch.float64, device=dev)
    ops.tap_gemm(A_hi, A_lo, W_hi, W_lo, D, a_phases=phases, a_rows=P, K=cin, b_slices=9, N=cout, M=P,
                 ldD=cout, Hp=Hp, Wp=Wp, shift=shift, phase=phase, bsel=bsel, stats=stats)
    yd = y.detach()
    assert rel(_from_pixel_cuda(D, n, cout, ho, ho), yd) < 5e-5
    assert rel(stats[0], yd.sum((0, 2, 3))) < 5e-5 * (n * ho * ho) ** 0.5
    assert rel(stats[1], (yd * yd).sum((0, 2, 3))) < 5e-5
    ring = D.view(n, Hp, Wp, cout).clone()
    ring[:, 1:ho + 1, 1:ho + 1] = 0
    assert float(ring.abs().max()) == 0.0
    # weight gradient: split-K workspace + fixed-order reduction, straight into OIHW
    G_hi, G_lo = split(_pixel_cuda(dy))
    ws = torch.empty(ops.wgrad_workspace_floats(P, cout, cin, 9), device=dev)
    assert ws.numel() >= 2 * 9 * cout * cin            # more than one split
    gw = torch.zeros(cout, cin, 3, 3, device=dev)
    ops.wgrad_gemm(G_hi, G_lo, A_hi, A_lo, gw, P=P, Cout=cout, x_phases=phases, Cin=cin, shift=shift,
                   phase=phase, bsel=bsel, oihw_taps=9, workspace=ws)
    assert rel(gw, wd.grad) < 5e-5
    # data gradient
    if stride == 1:
        dX = torch.full((P, cin), float('nan'), device=dev)
        ops.tap_gemm(G_hi, G_lo, Wt_hi, Wt_lo, dX, a_phases=1, a_rows=P, K=cout, b_slices=9, N=cin, M=P,
                     ldD=cin, Hp=Hp, Wp=Wp, shift=[-s for s in shift], phase=phase, bsel=bsel)
    else:
        dxp = torch.zeros(4, P, cin, device=dev)
        for ph in range(4):
            sh = [-shift[t] for t in range(9) if phase[t] == ph]
            bs = [bsel[t] for t in range(9) if phase[t] == ph]
            ops.tap_gemm(G_hi, G_lo, Wt_hi, Wt_lo, dxp[ph], a_phases=1, a_rows=P, K=cout, b_slices=9, N=cin,
                         M=P, ldD=cin, Hp=Hp, Wp=Wp, shift=sh, phase=[0] * len(sh), bsel=bs)
        dX = torch.empty(n * ops.padded(h) * ops.padded(h), cin, device=dev)
        ops.phase_unsplit(dxp, n, h, h, cin, dX)
    assert rel(_from_pixel_cuda(dX, n, cin, h, h), xd.grad) < 5e-5


@pytest.mark.parametrize('n,c,h', [(192, 64, 56), (192, 512, 7), (6, 128, 28)])
@pytest.mark.parametrize('with_residual', [False, True])
def test_tap_gemm_fused_bn_backward_epilogue(n, c, h, with_residual):
    """BwFuse epilogue of the data-gradient GEMM (csrc/gemm_tc.cu): D = dz = (dX + gb) * [act > 0] and
    the two BatchNorm-backward reductions sum(dz), sum(dz * xhat), against torch fp64 on identical
    inputs.  Shapes: the first and the last residual stage of BASELINE config 2 at full size."""
    from dmcnet_b200.engine import _taps_s1
    dev = torch.device('cuda')
    g = torch.Generator(device='cuda').manual_seed(12)
    dy = torch.randn(n, c, h, h, generator=g, device=dev)               # gradient of the conv OUTPUT
    w = torch.randn(c, c, 3, 3, generator=g, device=dev) * 0.05         # conv whose data gradient is taken
    Yn = torch.randn(n, c, h, h, generator=g, device=dev) * 1.5 + 0.3   # raw output of the BN'd conv upstream
    gbn = torch.randn(n, c, h, h, generator=g, device=dev) if with_residual else None
    mean = Yn.double().mean((0, 2, 3))
    var = Yn.double().var((0, 2, 3), unbiased=False)
    invstd = (var + 1e-5).rsqrt()
    act = torch.relu((Yn.double() - mean.view(1, c, 1, 1)) * invstd.view(1, c, 1, 1)
                     + 0.1 * torch.randn(n, c, h, h, generator=g, device=dev).double())
    # reference
    dX = torch.nn.grad.conv2d_input((n, c, h, h), w.double(), dy.double(), 1, 1)
    dz_ref = (dX + (gbn.double() if with_residual else 0.0)) * (act.to(torch.bfloat16).float() > 0)
    xhat = (Yn.double() - mean.view(1, c, 1, 1)) * invstd.view(1, c, 1, 1)
    s1_ref, s2_ref = dz_ref.sum((0, 2, 3)), (dz_ref * xhat).sum((0, 2, 3))
    # kernel
    Hp = Wp = ops.padded(h)
    P = n * Hp * Wp
    W_hi = torch.zeros(9, c, c, dtype=torch.bfloat16, device=dev)
    W_lo, Wt_hi, Wt_lo = (torch.zeros_like(W_hi) for _ in range(3))
    ops.weight_prep(w.contiguous(), c, c, 9, W_hi, W_lo, Wt_hi, Wt_lo)
    G_hi, G_lo = split(_pixel_cuda(dy))
    act_hi, _ = split(_pixel_cuda(act))
    Yp = _pixel_cuda(Yn)
    gb = _pixel_cuda(gbn) if with_residual else None
    shift, phase, bsel = _taps_s1(Wp)
    D = torch.full((P, c), float('nan'), device=dev)
    sums2 = torch.zeros(2, c, dtype=torch.float64, device=dev)
    ops.tap_gemm(G_hi, G_lo, Wt_hi, Wt_lo, D, a_phases=1, a_rows=P, K=c, b_slices=9, N=c, M=P, ldD=c,
                 Hp=Hp, Wp=Wp, shift=[-s for s in shift], phase=phase, bsel=bsel, stats=sums2,
                 bw=(Yp, act_hi, gb, mean.float().contiguous(), invstd.float().contiguous()))
    assert rel(_from_pixel_cuda(D, n, c, h, h), dz_ref) < 5e-5
    ring = D.view(n, Hp, Wp, c).clone()
    ring[:, 1:h + 1, 1:h + 1] = 0
    assert float(ring.abs().max()) == 0.0
    # sums of ~n*h*h terms of mixed sign: bar relative to the L2 size of the summands
    cnt = float(n * h * h)
    scale1 = float(dz_ref.abs().max()) * cnt ** 0.5
    assert float((sums2[0] - s1_ref).abs().max()) < 5e-5 * scale1
    scale2 = float((dz_ref * xhat).abs().max()) * cnt ** 0.5
    assert float((sums2[1] - s2_ref).abs().max()) < 5e-5 * scale2
    # and the BatchNorm backward completed from those sums equals autograd through batch_norm
    gamma = torch.rand(c, generator=g, device=dev) + 0.5
    Yg = Yn.double().requires_grad_(True)
    out = F.batch_norm(Yg, None, None, gamma.double(), None, True, 0.1, 1e-5)
    out.backward(dz_ref)
    Gh = torch.zeros(P, c, dtype=torch.bfloat16, device=dev)
    Gl = torch.zeros_like(Gh)
    dgam, dbet = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    ops.bn_bwd_apply(D, None, None, Yp, mean.float().contiguous(), invstd.float().contiguous(), gamma, sums2,
                     cnt, P, c, Hp, Wp, Gh, Gl, None, dgam, dbet)
    assert rel(_from_pixel_cuda(Gh.float() + Gl.float(), n, c, h, h), Yg.grad) < 5e-5


@pytest.mark.parametrize('n,c,h', [(192, 64, 56), (6, 128, 28), (3, 256, 14)])
def test_phase_unsplit_with_fused_mask_and_bn_backward_reductions(n, c, h):
    """dmc_phase_unsplit_reduce (csrc/pixelwise.cu) = dmc_phase_unsplit, then the ReLU mask and the two
    BatchNorm-backward reductions of dmc_bn_bwd_reduce / dmc_bn_bwd_apply's dz output, in one pass:
    dz bit-identical to the two-kernel sequence, sums to fp64 accuracy."""
    dev = torch.device('cuda')
    g = torch.Generator(device='cuda').manual_seed(21)
    Hp = Wp = ops.padded(h)
    Hq = Wq = ops.padded(h // 2)
    P, Pq = n * Hp * Wp, n * Hq * Wq
    dxp = torch.randn(4, Pq, c, generator=g, device=dev)
    Yn = torch.randn(n, c, h, h, generator=g, device=dev) * 1.5 + 0.3
    act = torch.relu(torch.randn(n, c, h, h, generator=g, device=dev))
    mean = Yn.double().mean((0, 2, 3)).float().contiguous()
    invstd = (Yn.double().var((0, 2, 3), unbiased=False) + 1e-5).rsqrt().float().contiguous()
    act_hi, _ = split(_pixel_cuda(act))
    Yp = _pixel_cuda(Yn)
    # two-kernel sequence
    dX = torch.empty(P, c, device=dev)
    ops.phase_unsplit(dxp, n, h, h, c, dX)
    dz_ref = _from_pixel_cuda(dX, n, c, h, h).double() * (_from_pixel_cuda(act_hi.float().view(P, c), n, c, h, h) > 0)
    xhat = (Yn.double() - mean.double().view(1, c, 1, 1)) * invstd.double().view(1, c, 1, 1)
    s_ref = torch.stack((dz_ref.sum((0, 2, 3)), (dz_ref * xhat).sum((0, 2, 3))))
    # fused
    out = torch.full((P, c), float('nan'), device=dev)
    sums2 = torch.zeros(2, c, dtype=torch.float64, device=dev)
    ops.phase_unsplit_reduce(dxp, n, h, h, c, act_hi, Yp, mean, invstd, out, sums2)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert torch.equal(_from_pixel_cuda(out, n, c, h, h).double(), dz_ref)
    ring = out.view(n, Hp, Wp, c)
    assert float(ring[:, 0].abs().max()) == 0.0 and float(ring[:, :, 0].abs().max()) == 0.0
    scale = dz_ref.abs().sum((0, 2, 3)).clamp_min(1e-30)
    assert float(((sums2 - s_ref).abs() / scale).max()) < 2e-6
