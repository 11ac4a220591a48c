"""CPU: host-side logic of the product path (no kernels are launched).

* the drop-in Model constructs the reference's state_dict (keys, shapes, init RNG order),
* the row-shifted "tap GEMM" formulation of 3x3 convolutions over the padded
  pixel-major layout (stride 1, stride 2 through four phases, and their data
  gradients) is checked in numpy against torch conv2d with the engine's own tap tables,
* product code refuses to run on CPU tensors (no fallback).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dmcnet_b200 import engine as E
from dmcnet_b200 import ops
from dmcnet_b200 import model as M
from oracle import dmc_oracle as O


@pytest.mark.parametrize('num_class,arch_d', [(51, None), (101, 'Discriminator3'), (51, 'Discriminator'),
                                              (51, 'Discriminator4')])
def test_model_state_matches_reference_constructor(num_class, arch_d):
    a = M.build_state(num_class, arch_d, seed=1)
    b = O.build_state(num_class, arch_d, seed=1)
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_model_api_surface():
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = M.Model(51, 3, 'mv', base_model='resnet18', arch_estimator='DenseNetTiny', use_databn=0,
                    gen_flow_or_delta=1)
        g = M.GANModel(51, 3, 'mv', base_model='resnet18', arch_estimator='DenseNetTiny', use_databn=1,
                       arch_d='Discriminator3')
        with pytest.raises(ValueError, match='Unknown base model'):
            M.Model(51, 3, 'mv', base_model='vgg16')
    assert m.crop_size == 224 and m.scale_size == 256 and m.num_segments == 3
    names = [n for n, _ in g.named_parameters()]
    assert any('base_model' in n for n in names) and any('gen_flow_model' in n for n in names)
    assert any('discriminator' in n for n in names) and 'data_bn.weight' in names
    assert not hasattr(m, 'discriminator')
    with pytest.raises(RuntimeError, match='CUDA'):
        m(torch.zeros(1, 3, 2, 224, 224), torch.zeros(1, 3, 3, 224, 224))


def _pad_pixel_major(x):          # [N,C,H,W] -> [N*Hp*Wp, C] with the (shared) zero ring
    n, c, h, w = x.shape
    p = np.zeros((n, ops.padded(h), ops.padded(w), c), np.float64)
    p[:, 1:h + 1, 1:w + 1, :] = x.transpose(0, 2, 3, 1)
    return p.reshape(-1, c)


def _tap_gemm(A, B, shift, phase, bsel, M_rows):
    """numpy model of dmc_tc_tap_gemm: A [phases][rows][K], B [slices][N][K]."""
    rows = A.shape[1]
    out = np.zeros((M_rows, B.shape[1]))
    q = np.arange(M_rows)
    for s, p, b in zip(shift, phase, bsel):
        idx = q + s
        ok = (idx >= 0) & (idx < rows)
        a = A[p][np.clip(idx, 0, rows - 1)] * ok[:, None]
        out += a @ B[b].T
    return out


def _interior(n, h, w, c, flat):
    return flat.reshape(n, ops.padded(h), ops.padded(w), c)[:, 1:h + 1, 1:w + 1, :].transpose(0, 3, 1, 2)


def _phase_split(x):               # [N,C,H,W] -> [4][N*(H/2+2)*(W/2+2), C]
    return np.stack([_pad_pixel_major(x[:, :, ph::2, pw::2]) for ph in (0, 1) for pw in (0, 1)])


def test_tap_tables_stride1_and_stride2_match_conv2d():
    rng = np.random.default_rng(0)
    n, ci, co, h, w = 2, 5, 4, 8, 6
    x = rng.standard_normal((n, ci, h, w))
    wt = rng.standard_normal((co, ci, 3, 3))
    Wg = wt.transpose(2, 3, 0, 1).reshape(9, co, ci)               # [tap][co][ci]
    # stride 1
    shift, phase, bsel = E._taps_s1(ops.padded(w))
    A = _pad_pixel_major(x)[None]
    y = _interior(n, h, w, co, _tap_gemm(A, Wg, shift, phase, bsel, A.shape[1]))
    ref = F.conv2d(torch.tensor(x), torch.tensor(wt), None, 1, 1).numpy()
    np.testing.assert_allclose(y, ref, atol=1e-10)
    # stride 2 through the four input phases stored in the OUTPUT geometry
    shift, phase, bsel = E._taps_s2(ops.padded(w // 2))
    A2 = _phase_split(x)
    y2 = _interior(n, h // 2, w // 2, co, _tap_gemm(A2, Wg, shift, phase, bsel, A2.shape[1]))
    ref2 = F.conv2d(torch.tensor(x), torch.tensor(wt), None, 2, 1).numpy()
    np.testing.assert_allclose(y2, ref2, atol=1e-10)


def test_tap_tables_data_gradients():
    """dgrad = tap GEMM with negated shifts and transposed weights (stride 1), and one
    GEMM per input phase with the matching tap subset (stride 2), as engine._cls_backward."""
    rng = np.random.default_rng(1)
    n, ci, co, h, w = 1, 3, 4, 6, 8
    wt = rng.standard_normal((co, ci, 3, 3))
    Wt = wt.transpose(2, 3, 1, 0).reshape(9, ci, co)               # [tap][ci][co]
    x = torch.tensor(rng.standard_normal((n, ci, h, w)), requires_grad=True)
    for stride in (1, 2):
        y = F.conv2d(x, torch.tensor(wt), None, stride, 1)
        dy = rng.standard_normal(tuple(y.shape))
        (gx,) = torch.autograd.grad(y, x, torch.tensor(dy))
        G = _pad_pixel_major(dy)[None]
        ho, wo = h // stride, w // stride
        if stride == 1:
            shift, phase, bsel = E._taps_s1(ops.padded(w))
            dx = _interior(n, h, w, ci, _tap_gemm(G, Wt, [-s for s in shift], phase, bsel, G.shape[1]))
        else:
            fs, fp, fb = E._taps_s2(ops.padded(wo))
            dx = np.zeros((n, ci, h, w))
            for ph in range(4):
                sh = [-fs[t] for t in range(9) if fp[t] == ph]
                bs = [fb[t] for t in range(9) if fp[t] == ph]
                d = _interior(n, ho, wo, ci, _tap_gemm(G, Wt, sh, [0] * len(sh), bs, G.shape[1]))
                dx[:, :, ph // 2::2, ph % 2::2] = d
        np.testing.assert_allclose(dx, gx.numpy(), atol=1e-10)


def test_generator_channel_layout():
    """New channels are prepended (code/dmcnet/model.py:186-194): layer k reads a
    contiguous channel range that ends with mv(2)+res(3)."""
    outs, off = [], 28
    for g in E.GEN_GROWTH:
        off -= g
        outs.append(off)
    assert outs == [20, 12, 6, 2, 0]
    cins = [33 - (o + g) for o, g in zip(outs, E.GEN_GROWTH)]
    assert cins == [5, 13, 21, 27, 31]


def test_disc_block_tables_match_oracle():
    for a in ('Discriminator', 'Discriminator2', 'Discriminator3', 'Discriminator4', 'Discriminator5'):
        assert E.disc_blocks(a) == O.disc_blocks(a)


def _s2_tap(p, a):                 # mirrors s2_tap() in csrc/dense_conv.cu
    return (1 if a == 1 else -1) if p == 0 else (0 if a == 0 else (2 if a == 1 else -1))


def _s2d(x):                       # [N,C,H,W] -> [N,4C,H/2,W/2], channel = (pr*2+pc)*C + c
    return torch.cat([x[:, :, pr::2, pc::2] for pr in (0, 1) for pc in (0, 1)], 1)


def _s2_weight(w):                 # [Co,Ci,3,3] -> [Co,4Ci,3,3]
    co, ci = w.shape[:2]
    w3 = torch.zeros(co, 4 * ci, 3, 3, dtype=w.dtype)
    for p in range(4):
        for a in range(3):
            for b in range(3):
                r, s_ = _s2_tap(p >> 1, a), _s2_tap(p & 1, b)
                if r >= 0 and s_ >= 0:
                    w3[:, p * ci:(p + 1) * ci, a, b] = w[:, :, r, s_]
    return w3


def test_stride2_conv_equals_stride1_conv_on_space_to_depth():
    """The identity behind the discriminator's stride-2 backward (engine._disc_backward):
    conv3x3/2(x, w) == conv3x3/1(s2d(x), W3), and the gradients map back by gather / depth-to-space."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 8, 16, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(5, 3, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, None, 2, 1)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    gx, gw = torch.autograd.grad(y, (x, w), dy)
    S = _s2d(x.detach()).requires_grad_(True)
    W3 = _s2_weight(w.detach()).requires_grad_(True)
    y3 = F.conv2d(S, W3, None, 1, 1)
    assert torch.allclose(y3, y.detach(), atol=1e-12)
    gS, gW3 = torch.autograd.grad(y3, (S, W3), dy)
    # weight gradient: every OIHW element has exactly one source in W3
    gw_back = torch.zeros_like(gw)
    hits = torch.zeros(3, 3)
    for p in range(4):
        for a in range(3):
            for b in range(3):
                r, s_ = _s2_tap(p >> 1, a), _s2_tap(p & 1, b)
                if r >= 0 and s_ >= 0:
                    gw_back[:, :, r, s_] += gW3[:, p * 3:(p + 1) * 3, a, b]
                    hits[r, s_] += 1
    assert torch.equal(hits, torch.ones(3, 3))
    assert torch.allclose(gw_back, gw, atol=1e-12)
    # data gradient: depth-to-space of dS
    gx_back = torch.zeros_like(gx)
    for p in range(4):
        gx_back[:, :, p >> 1::2, p & 1::2] = gS[:, p * 3:(p + 1) * 3]
    assert torch.allclose(gx_back, gx, atol=1e-12)


def test_stem_im2col_operand_layout():
    """Operand layout of the tensor-core stem conv (csrc/stem_tc.cu): K = 128 = 8 kernel rows x
    2 channels x 8 columns, k = r*16 + ci*8 + s; the 16-byte chunk (r, ci) of output pixel (y, x)
    is the 8 consecutive input pixels of row 2y+r-3 starting at column 2x-3; the weight operand is
    zero at r = 7 and s = 7.  A @ Wb^T must equal the 7x7 stride-2 pad-3 convolution."""
    rng = np.random.default_rng(3)
    n, H, W = 1, 12, 16
    x = rng.standard_normal((n, 2, H, W))
    w = rng.standard_normal((64, 2, 7, 7))
    Ho, Wo = H // 2, W // 2
    xp = np.zeros((n, 2, H + 8, W + 8))                 # staging: image column c lives at c + 3
    xp[:, :, 3:3 + H, 3:3 + W] = x
    A = np.zeros((n, Ho, Wo, 128))
    for y in range(Ho):
        for xo in range(Wo):
            for r in range(7):
                for ci in range(2):
                    A[0, y, xo, r * 16 + ci * 8:r * 16 + ci * 8 + 8] = xp[0, ci, 2 * y + r, 2 * xo:2 * xo + 8]
    Wb = np.zeros((64, 128))
    for r in range(7):
        for ci in range(2):
            Wb[:, r * 16 + ci * 8:r * 16 + ci * 8 + 7] = w[:, ci, r, :]
    y_gemm = (A.reshape(-1, 128) @ Wb.T).reshape(n, Ho, Wo, 64).transpose(0, 3, 1, 2)
    ref = F.conv2d(torch.tensor(x), torch.tensor(w), None, 2, 3).numpy()
    np.testing.assert_allclose(y_gemm, ref, atol=1e-10)



def _cpu_engine_tables(growth, H=16, W=16, N=2):
    """The engine's parameter table and generator plan built on the CPU (no kernels): the host
    logic under test is the channel layout and the combined-weight table of _alloc_generator."""
    e = object.__new__(E.DmcEngine)
    e.device = torch.device('cpu')
    e.num_class, e.S, e.N = 51, 3, N
    e.gan, e.arch_d, e.gen_flow_or_delta, e.H, e.W = False, None, 1, H, W
    e.gen_growth = tuple(growth)
    e._build_param_table()
    e._alloc_generator()
    return e


@pytest.mark.parametrize('growth', [(8, 8, 6, 4, 2), (32, 32, 24, 16, 8), (12, 4, 8, 2, 6)])
def test_dense_block_gradient_plan_matches_autograd(growth):
    """Backward plan of the dense generator (engine._gen_backward): the gradient of dense slice k is
    ONE 3x3 convolution over [d gen_flow | d new_{L-1} .. d new_{k+1}] with weights gathered from the
    flipped weights of every later layer (numpy model of dense_dgrad_weights_kernel driven by the
    engine's own table), times LeakyReLU'.  Checked against torch autograd for the Tiny table, the
    DenseNetSmall table and an irregular one."""
    e = _cpu_engine_tables(growth)
    N, H, W, L = e.N, e.H, e.W, len(growth)
    g = torch.Generator().manual_seed(sum(growth))
    e.params.copy_(torch.randn(e.params.shape, generator=g) * 0.1)
    x = torch.randn(N, 5, H, W, generator=g)
    # reference: EstimatorDense forward (code/dmcnet/model.py:186-194) with autograd
    pre, cur = [], x
    for k in range(L):
        w = e.param_view('gen_flow_model.conv_%d.0.weight' % k).clone().requires_grad_(True)
        z = F.conv2d(cur, w, e.param_view('gen_flow_model.conv_%d.0.bias' % k), 1, 1)
        z.retain_grad()
        pre.append(z)
        cur = torch.cat((F.leaky_relu(z, 0.1), cur), 1)
    out = F.conv2d(cur, e.param_view('gen_flow_model.predict_flow.weight'),
                   e.param_view('gen_flow_model.predict_flow.bias'), 1, 1)
    R = torch.randn(out.shape, generator=g)
    (out * R).sum().backward()
    # the engine's buffers: X = [new_{L-1} .. new_0 | mv res], dD = [d gen_flow | d new_{L-1} .. d new_0]
    X = cur.detach()
    assert X.shape[1] == e.gen_ctot
    for k in range(L):
        oo = e.gen_out_off[k]
        assert torch.equal(X[:, oo:oo + growth[k]], F.leaky_relu(pre[k].detach(), 0.1))
    dD = torch.zeros(e.dD.shape)
    dD[:, 0:2] = R
    tab = e.gen_wc_table
    assert tab[0] == L
    params = e.params.numpy()
    for idx, k in enumerate(reversed(range(L))):
        row = tab[1 + idx * 34: 1 + (idx + 1) * 34]
        out_off, gk, cin_s, nseg = row[:4]
        assert (out_off, gk, cin_s) == (e.gen_wc_off[idx], growth[k], 2 + e.gen_out_off[k])
        wc = np.zeros((gk, cin_s, 9), np.float32)
        for s in range(nseg):
            c0, cnt, w_off, cin_j, ci_off = row[4 + 5 * s: 9 + 5 * s]
            for c in range(c0, c0 + cnt):
                for ci in range(gk):
                    base = w_off + ((c - c0) * cin_j + ci_off + ci) * 9
                    wc[ci, c, :] = params[base:base + 9][::-1]
        d_post = F.conv2d(dD[:, :cin_s], torch.from_numpy(wc).view(gk, cin_s, 3, 3), None, 1, 1)
        oo = e.gen_out_off[k]
        act = X[:, oo:oo + gk]
        d_pre = d_post * torch.where(act > 0, torch.ones(()), torch.full((), 0.1))
        dD[:, 2 + oo:2 + oo + gk] = d_pre
        ref = pre[k].grad
        assert float((d_pre - ref).abs().max()) <= 2e-5 * float(ref.abs().max()), (growth, k)


@pytest.mark.parametrize('arch,att,ds', [('DenseNetTiny', 0, 0), ('DenseNetSmall', 0, 0), ('DenseNet', 0, 4),
                                         ('DenseNetTinyEarlyFusionSum', 0, 0), ('DenseNetTinyEarlyFusionStack', 0, 4),
                                         ('ContextNetwork', 0, 0), ('ContextNetwork', 1, 0), ('ContextNetwork', 0, 4),
                                         ('ContextNetwork', 1, 4)])
def test_engine_parameter_table_matches_the_reference_state_dict_for_every_generator(arch, att, ds):
    """Keys, order and shapes of the engine's parameter / buffer table against the state_dict the
    oracle builds (itself bit-pinned on the reference constructors, oracle/pin_against_reference.py)."""
    from sim_engine import SimEngine
    from oracle import dmc_oracle as O
    sd = O.build_state(11, 'Discriminator', seed=1, arch_estimator=arch, att=att, gen_flow_ds_factor=ds)
    growth = O.DENSE_GROWTH.get(arch, O.GEN_TINY_GROWTH)
    eng = SimEngine(11, 3, 3, gan=True, arch_d='Discriminator', height=32, width=32, gen_growth=growth,
                    arch_estimator=arch, att=att, gen_flow_ds_factor=ds)
    keys = eng.state_keys()
    want = [k for k in sd if not k.startswith('discriminator.adv_layer')]       # fc_in depends on the frame size
    assert [k for k in keys if not k.startswith('discriminator.adv_layer')] == want
    for k in want:
        shp = eng.specs[k] if k in eng.specs else tuple(eng.buffers[k].shape)
        assert tuple(shp) == tuple(sd[k].shape), k
