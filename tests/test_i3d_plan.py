"""Host logic of the I3D execution plan (dmcnet_b200/i3d_engine.py) on the CPU: parameter table vs the
reference state_dict, and the GEMM formulation -- padded column layout of the concatenated maps, weight
gather tables, 27-tap row shifts on the shared-ring layout, stem im2col column order -- emulated with torch
matmuls and compared with the oracle's F.conv3d blocks."""
import torch
import torch.nn.functional as F
import pytest

from dmcnet_b200 import i3d_engine as E
from oracle import i3d_oracle as O


def cpu_engine(num_class=51, clips=1, clip_len=16):
    eng = object.__new__(E.I3DEngine)
    eng.device = torch.device('cpu')
    eng.num_class, eng.clips, eng.clip_len = num_class, clips, clip_len
    eng.has_gen, eng.gan, eng.arch_d = True, False, None
    eng.gen_growth = E.GEN_TABLE['DenseNetTiny']
    eng.H = eng.W = 224
    eng._share_from = None
    eng._plan_trunk(clips, clip_len, 224, 224)
    eng._build_param_table()
    return eng


def test_parameter_table_matches_reference_state_dict():
    eng = cpu_engine()
    sd = O.build_state(51, 'DenseNetTiny', seed=1)
    assert eng.state_keys() == list(sd.keys())
    for k, v in sd.items():
        shp = tuple(v.shape)
        got = tuple(eng.specs[k]) if k in eng.specs else tuple(eng.buffers[k].shape)
        assert got == shp, k
    eng.load_state(sd)
    out = eng.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k], v), k
    # spans do not overlap: every parameter element is owned by exactly one key
    owner = torch.zeros(eng.total, dtype=torch.int32)
    for k in eng.specs:
        o = eng.offsets[k]
        owner[o:o + eng.numel(k)] += 1
    assert int(owner.max()) == 1
    assert [g.thw for g in eng.geos] == [(8, 112, 112), (8, 56, 56), (8, 28, 28), (4, 14, 14), (2, 7, 7)]


def to_ring(x):          # [B,C,T,H,W] -> padded pixel-major [B*(T+1)*(H+1)*(W+1), C]
    b, c, t, h, w = x.shape
    p = torch.zeros(b, t + 1, h + 1, w + 1, c, dtype=x.dtype)
    p[:, 1:, 1:, 1:, :] = x.permute(0, 2, 3, 4, 1)
    return p.reshape(-1, c)


def from_ring(flat, b, c, t, h, w):
    return flat.reshape(b, t + 1, h + 1, w + 1, -1)[:, 1:, 1:, 1:, :c].permute(0, 4, 1, 2, 3)


def ring_mask(b, t, h, w):
    m = torch.zeros(b, t + 1, h + 1, w + 1, 1)
    m[:, 1:, 1:, 1:] = 1
    return m.reshape(-1, 1)


def tap_gemm(A, Wg, shifts, bsel, mask):
    """out[q] = sum_t A[q + shift_t] . Wg[bsel_t]^T on the flat layout (reads past the ends are zero)."""
    P, K = A.shape
    pad = max(abs(s) for s in shifts) + 1
    Ap = torch.zeros(P + 2 * pad, K, dtype=A.dtype)
    Ap[pad:pad + P] = A
    out = torch.zeros(P, Wg.shape[1], dtype=A.dtype)
    for s, b in zip(shifts, bsel):
        out += Ap[pad + s:pad + s + P] @ Wg[b].t()
    return out * mask


def gemm_weights(eng, units, params):
    n_total = sum(u.cop for u in units)
    T, Kp = units[0].T, units[0].Kp
    gmap = torch.full((T * n_total * Kp,), -1, dtype=torch.int32)
    n0 = 0
    for u in units:
        inv = torch.full((u.cout * u.cin * (343 if u.k == 7 else u.T),), -1, dtype=torch.int32)
        eng._weight_tables(u, n_total, n0, gmap, inv)
        # inverse table: every weight element sits at exactly one operand position
        assert int((inv < 0).sum()) == 0
        assert torch.equal(gmap[inv.long()].long(), eng.offsets[u.name + '.conv3d.weight'] + torch.arange(inv.numel()))
        n0 += u.cop
    W = torch.where(gmap >= 0, params[gmap.clamp_min(0).long()], torch.zeros(()))
    return W.view(T, n_total, Kp)


@pytest.mark.parametrize('block', [0, 2, 5])
def test_inception_block_as_column_slice_gemms(block):
    torch.manual_seed(3)
    eng = cpu_engine()
    sd = O.build_state(51, 'DenseNetTiny', seed=2)
    eng.load_state(sd)
    params = eng.params.double()
    M = eng.mixed[block]
    name, cin = M['name'], E.MIXED[block][1]
    b, t, h, w = 1, 3, 4, 5
    x = torch.relu(torch.randn(b, cin, t, h, w, dtype=torch.float64))
    st = {k: v.double() if v.is_floating_point() else v.clone() for k, v in sd.items()}
    ref = O._mixed(st, name, x, train=True)
    # the same block as GEMMs on the padded-column layout
    geo = E._Geo3(b, t, h, w)
    mask = ring_mask(b, t, h, w).double()
    cols = torch.tensor(M['b0'].in_cols)
    A = torch.zeros(geo.P, M['Kin'], dtype=torch.float64)
    A[:, cols] = to_ring(x).double()

    def bn_relu(Y, gr):
        n = geo.count
        mean = Y.sum(0) / n
        var = (Y * Y).sum(0) / n - mean * mean
        gamma = params[gr.gamma_off:gr.gamma_off + gr.width]
        beta = params[gr.beta_off:gr.beta_off + gr.width]
        return torch.relu((Y - mean) * (var + 1e-5).rsqrt() * gamma + beta) * mask

    s1, b1 = geo.taps(1)
    s3, b3 = geo.taps(3)
    Ymid = tap_gemm(A, gemm_weights(eng, [M['b1a'], M['b2a']], params), s1, b1, mask)
    Amid = bn_relu(Ymid, M['mid'])
    Ycat = torch.zeros(geo.P, M['cat'].width, dtype=torch.float64)
    u = M['b0']
    Ycat[:, u.col0:u.col0 + u.cop] = tap_gemm(A, gemm_weights(eng, [u], params), s1, b1, mask)
    for ua, ub in ((M['b1a'], M['b1b']), (M['b2a'], M['b2b'])):
        Ycat[:, ub.col0:ub.col0 + ub.cop] = tap_gemm(Amid[:, ua.col0:ua.col0 + ua.cop], gemm_weights(eng, [ub], params),
                                                     s3, b3, mask)
    xp = F.max_pool3d(F.pad(x, (1, 1, 1, 1, 1, 1)), 3, 1)
    Ap = torch.zeros_like(A)
    Ap[:, cols] = to_ring(xp).double()
    u = M['b3']
    Ycat[:, u.col0:u.col0 + u.cop] = tap_gemm(Ap, gemm_weights(eng, [u], params), s1, b1, mask)
    out = bn_relu(Ycat, M['cat'])
    # padding columns stay exactly zero; real columns are the reference's channels in concat order
    nxt_cols = []
    for u in (M['b0'], M['b1b'], M['b2b'], M['b3']):
        nxt_cols += [u.col0 + c for c in range(u.cout)]
    got = from_ring(out[:, nxt_cols], b, len(nxt_cols), t, h, w)
    assert float((got - ref).abs().max()) < 1e-9
    pad_cols = sorted(set(range(M['cat'].width)) - set(nxt_cols))
    if pad_cols:
        assert float(out[:, pad_cols].abs().max()) == 0.0
    if block + 1 < len(eng.mixed):
        assert eng.mixed[block + 1]['b0'].in_cols == nxt_cols


def test_stem_as_seven_temporal_taps_over_frame_patches():
    """Conv3d(2, 64, 7, stride 2), TF "SAME" (2 zeros in front, 3 behind): per-frame 7x7x2 patches at the
    output's spatial positions, even / odd frames as two phases, one tap per temporal kernel index."""
    torch.manual_seed(4)
    eng = cpu_engine()
    sd = O.build_state(51, 'DenseNetTiny', seed=2)
    eng.load_state(sd)
    params = eng.params.double()
    clips, T, H, W = 2, 8, 6, 8
    x = torch.randn(clips, 2, T, H, W, dtype=torch.float64)
    w = sd['conv3d_1a_7x7.conv3d.weight'].double()
    ref = F.conv3d(F.pad(x, O.tf_same_pad((7, 7, 7), (2, 2, 2))), w, None, 2)
    Wg = gemm_weights(eng, [eng.stem], params)                        # [7][64][128]
    geo = E._Geo3(clips, T // 2, H // 2, W // 2, t_hi=1)
    A2 = torch.zeros(2, clips, geo.Tp, geo.Hp, geo.Wp, E.STEM_KP, dtype=torch.float64)
    for n in range(clips):
        for t in range(T):
            for ho in range(geo.H):
                for wo in range(geo.W):
                    for kh in range(7):
                        for kw in range(7):
                            h, ww = 2 * ho + kh - 2, 2 * wo + kw - 2
                            if 0 <= h < H and 0 <= ww < W:
                                A2[t % 2, n, t // 2 + 1, ho + 1, wo + 1, (kh * 7 + kw) * 2:(kh * 7 + kw) * 2 + 2] = x[n, :, t, h, ww]
    fr = geo.Hp * geo.Wp
    shift = [((kt // 2 - 1) if kt % 2 == 0 else (kt - 3) // 2) * fr for kt in range(7)]
    P = geo.P
    out = torch.zeros(P, 64, dtype=torch.float64)
    pad = 3 * fr
    for kt in range(7):
        Ap = torch.zeros(P + 2 * pad, E.STEM_KP, dtype=torch.float64)
        Ap[pad:pad + P] = A2[kt % 2].reshape(P, E.STEM_KP)
        out += Ap[pad + shift[kt]:pad + shift[kt] + P] @ Wg[kt].t()
    got = out.reshape(clips, geo.Tp, geo.Hp, geo.Wp, 64)[:, 1:1 + geo.T, 1:, 1:].permute(0, 4, 1, 2, 3)
    assert tuple(ref.shape[2:]) == geo.thw
    assert float((got - ref).abs().max()) < 1e-9
