"""The round-2 kernels behind the whole-subgraph autograd Functions (fused_stack.py), each against a plain torch restatement
through the C ABI: multi-copy dropout and its merged backward, the question-word prologue (embedding + dropout + tanh) and its
scatter-add backward, two-stream view attention, grouped weight casts, LSTM bias / gradient packing, the de-interleaving grouped
column sum, synchronised-BatchNorm passes, the end-of-step finaliser; plus the two fused Functions against the module-by-module
path they replace (same kernels underneath => tight tolerance) and the dynamic tile schedule of the LSTM kernels under
concurrent launches."""
import pytest
import torch

import dualvgr_oracle as orc

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ops():
    import dualvgr_videoqa_b200.ops as ops
    import dualvgr_videoqa_b200._lib as L
    L.lib.dvgr_set_seed_offset(None)
    return ops


def test_dropout_multi_and_merged_backward(ops):
    g = torch.Generator().manual_seed(3)
    M, D = 200, 768
    x = [torch.randn((M, D), generator=g).to(BF16).cuda() for _ in range(2)]
    seed, p = 1234, 0.15
    outs = [torch.empty_like(x[0]) for _ in range(4)]
    ops.dropout_multi([x[0], x[0], x[1], x[1]], outs, [7, 8, 9, 10], p, seed)
    for i in range(4):
        ref = ops.dropout_raw(x[i // 2], p, seed, 7 + i)
        assert torch.equal(outs[i], ref)
    assert not torch.equal(outs[0], outs[1])                      # different streams -> different masks
    keep = float((outs[0] != 0).float().mean())
    assert abs(keep - (1 - p)) < 0.01
    # backward: dX[s] = base[s] + sum_j mask_j * dxt[2s + j]
    dxt = [torch.randn((M, D), generator=g).to(BF16).cuda() for _ in range(4)]
    base = [torch.randn((M, D), generator=g).to(BF16).cuda() for _ in range(2)]
    out = [torch.empty_like(base[0]) for _ in range(2)]
    ops.gat_input_bwd(dxt, [7, 8, 9, 10], 2, base, out, p, seed)
    for s in range(2):
        ref = base[s].float() + ops.dropout_raw(dxt[2 * s], p, seed, 7 + 2 * s).float() + \
            ops.dropout_raw(dxt[2 * s + 1], p, seed, 8 + 2 * s).float()
        assert rel(out[s], ref) < 4e-3
    # p = 0 and no base: plain sum
    ops.gat_input_bwd(dxt, [0, 0, 0, 0], 2, [None, None], out, 0.0, seed)
    assert rel(out[1], dxt[2].float() + dxt[3].float()) < 4e-3


def test_question_word_prologue_and_scatter_backward(ops):
    g = torch.Generator().manual_seed(5)
    B, L, V, W, Wp = 9, 7, 40, 300, 304
    table = (torch.rand((V, W), generator=g) * 2 - 1).cuda()
    tok = torch.randint(0, V, (B, L), generator=g).cuda()
    words, x_tm = ops.embed_fwd(tok, table, Wp)
    ref = torch.tanh(table[tok])
    assert rel(words[..., :W], ref) < 4e-3 and float(words[..., W:].abs().max()) == 0
    assert torch.equal(x_tm, words.transpose(0, 1).contiguous())
    # dropout: same mask in forward and backward, keep rate
    wd, _ = ops.embed_fwd(tok, table, Wp, 0.15, 99, 3)
    kept = (wd[..., :W] != 0) | (ref.abs() < 1e-3)
    assert abs(float(kept.float().mean()) - 0.85) < 0.02
    # backward: dtable = scatter-add of (d_words + d_x^T) * (1 - w^2) * mask
    dwo = torch.randn((B, L, Wp), generator=g).to(BF16).cuda()
    dxt = torch.randn((L, B, Wp), generator=g).to(BF16).cuda()
    dt = torch.zeros_like(table)
    ops.embed_bwd(tok, wd, dwo, dxt, W, dt, 0.15, 99, 3)
    mask = (wd[..., :W] != 0).float() / 0.85
    d = (dwo.float() + dxt.float().transpose(0, 1))[..., :W] * (1 - wd[..., :W].float() ** 2) * mask
    ref_dt = torch.zeros_like(table).index_put_((tok.reshape(-1),), d.reshape(-1, W), accumulate=True)
    assert rel(dt, ref_dt) < 1e-5


def test_two_stream_view_attention_equals_two_single_calls(ops):
    g = torch.Generator().manual_seed(7)
    M, D = 301, 768
    hidden = torch.tanh(torch.randn((2, 2, M, D), generator=g)).to(BF16).cuda()
    z = torch.randn((2, 2, M, D), generator=g).to(BF16).cuda()
    x = torch.randn((2, M, D), generator=g).to(BF16).cuda()
    w2 = (torch.randn((2, D), generator=g) * 0.05).cuda()
    xn, em, beta = ops.view_attn_fwd_multi(hidden, z, x, w2)
    dxn = torch.randn((2, M, D), generator=g).to(BF16).cuda()
    dem = torch.randn((2, M, D), generator=g).to(BF16).cuda()
    dz, dhid, part = ops.view_attn_bwd_multi(dxn, dem, hidden, z, w2, beta)
    for s in range(2):
        xn1, em1, b1 = ops.view_attn_fwd(hidden[s], z[s], x[s], w2[s])
        assert torch.equal(xn[s], xn1) and torch.equal(em[s], em1) and torch.equal(beta[s], b1)
        dz1, dh1, dw1 = ops.view_attn_bwd(dxn[s], dem[s], hidden[s], z[s], w2[s], b1)
        assert torch.equal(dz[s], dz1) and torch.equal(dhid[s], dh1)
        assert rel(ops.colsum(part[s]), dw1) < 1e-5


def test_grouped_casts_lstm_packing_and_permuted_column_sum(ops):
    g = torch.Generator().manual_seed(9)
    H, K = 64, 72
    ws = [torch.randn((4 * H, K), generator=g).cuda() for _ in range(4)]
    buf = torch.empty((16 * H, 80), dtype=BF16, device="cuda")
    ops.cast_rows_grouped(ws, buf, out_cols=80, lstm_H=H)
    for d in range(4):
        assert torch.equal(buf[d * 4 * H:(d + 1) * 4 * H], ops.cast_rows(ws[d], out_cols=80, lstm_H=H))
    bi = [torch.randn(4 * H, generator=g).cuda() for _ in range(4)]
    bh = [torch.randn(4 * H, generator=g).cuda() for _ in range(4)]
    packed = ops.lstm_pack_bias(bi, bh, H)
    for d in range(4):
        ref = (bi[d] + bh[d]).view(4, H).t().reshape(-1)
        assert torch.equal(packed[d * 4 * H:(d + 1) * 4 * H], ref)
    # gradient packing for the whole-sequence backward
    S, T, D = 45, 6, 4
    d_seq = torch.randn((S, T, 2 * H), generator=g).to(BF16).cuda()
    d_last = torch.randn((S, 2 * H), generator=g).to(BF16).cuda()
    wide = torch.zeros((S * T, 4 * H), dtype=BF16, device="cuda")
    wide[:, :2 * H] = d_seq.view(S * T, 2 * H)
    dh_seq, dh_last = ops.lstm_pack_dh(wide[:, :2 * H].view(S, T, 2 * H), 2, d_last, 2, S, T, D, H)
    RB, UG = (S + 31) // 32, H // 8
    full = torch.zeros((RB * 32, T, D * H), dtype=BF16, device="cuda")
    full[:S, :, :2 * H] = d_seq
    ref = full.view(RB, 32, T, D, UG, 8).permute(2, 3, 0, 4, 1, 5).contiguous()
    assert torch.equal(dh_seq, ref)
    ref_last = torch.zeros((S, D * H), dtype=BF16, device="cuda")
    ref_last[:, 2 * H:] = d_last
    assert torch.equal(dh_last, ref_last)
    # grouped column sum with the LSTM de-interleave and a second target
    dg = torch.randn((500, 2 * 4 * H), generator=g).to(BF16).cuda()
    t1, t2 = torch.zeros(4 * H, device="cuda"), torch.ones(4 * H, device="cuda")
    ops.colsum_enqueue(dg[:, 4 * H:], t1, perm_H=H, out2=t2)
    ops.flush_wgrads()
    ref = dg[:, 4 * H:].float().sum(0).view(H, 4).t().reshape(-1)
    assert rel(t1, ref) < 1e-5 and rel(t2 - 1, ref) < 1e-5


def test_sync_batchnorm_passes_and_finaliser(ops):
    g = torch.Generator().manual_seed(11)
    B, D = 48, 768
    x = torch.randn((B, D), generator=g).cuda() * 0.7 + 0.2
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=g)).cuda(), (0.1 * torch.randn(D, generator=g)).cuda()
    rm, rv = torch.zeros(D, device="cuda"), torch.ones(D, device="cuda")
    y_ref, m_ref, r_ref = ops.bn_fwd(x, gamma, beta, rm.clone(), rv.clone(), True)
    # "two ranks": halves of the batch; the summed statistics must reproduce the single-batch normalisation
    stats = ops.bn_stats(x[:24]) + ops.bn_stats(x[24:])
    assert rel(stats[0], x.sum(0)) < 1e-5 and rel(stats[1], (x * x).sum(0)) < 1e-5
    rm2, rv2 = rm.clone(), rv.clone()
    y0, m0, r0 = ops.bn_fwd(x[:24].contiguous(), gamma, beta, rm2, rv2, True, ext_stats=stats, Btot=B)
    y1, _, _ = ops.bn_fwd(x[24:].contiguous(), gamma, beta, rm.clone(), rv.clone(), True, ext_stats=stats, Btot=B)
    assert rel(torch.cat([y0, y1]), y_ref) < 5e-3 and rel(m0, m_ref) < 1e-5 and rel(r0, r_ref) < 1e-4
    rm3, rv3 = rm.clone(), rv.clone()
    ops.bn_fwd(x, gamma, beta, rm3, rv3, True)
    assert rel(rm2, rm3) < 1e-5 and rel(rv2, rv3) < 1e-4
    dy = torch.randn((B, D), generator=g).to(BF16).cuda()
    dx_ref, dg_ref, db_ref = ops.bn_bwd(dy, x, gamma, m_ref, r_ref, True)
    halves = [(dy[:24].contiguous(), x[:24].contiguous()), (dy[24:].contiguous(), x[24:].contiguous())]
    loc = [ops.bn_bwd(d, xx, gamma, m_ref, r_ref, True, stats_only=True) for d, xx in halves]
    sums = torch.stack([loc[0][2] + loc[1][2], loc[0][1] + loc[1][1]])
    assert rel(sums[0], db_ref) < 1e-5 and rel(sums[1], dg_ref) < 1e-5
    dx = torch.cat([ops.bn_bwd(d, xx, gamma, m_ref, r_ref, True, ext_sums=sums, Btot=B)[0] for d, xx in halves])
    assert rel(dx, dx_ref) < 1e-5
    # finaliser: total, terms, and the NaN poison when a sticky error word is set
    ce = torch.tensor([1.25], device="cuda")
    parts = torch.rand((3, 10, 3), generator=g).cuda()
    ok, bad = torch.zeros(1, dtype=torch.int32, device="cuda"), torch.ones(1, dtype=torch.int32, device="cuda")
    out = ops.finalize_loss(ce, parts, [ok, ok])
    s = parts.view(-1, 3).sum(0)
    assert abs(float(out[0]) - float(1.25 + s.sum())) < 1e-4 and abs(float(out[1]) - float(s[0])) < 1e-4
    assert abs(float(out[2]) - float(s[1] + s[2])) < 1e-4 and float(out[3]) == 0
    out = ops.finalize_loss(ce, parts, [ok, bad])
    assert bool(torch.isnan(out[0])) and float(out[3]) == 1
    out = ops.finalize_loss(ce, None, [])
    assert abs(float(out[0]) - 1.25) < 1e-6


def _model(cfg, p_zero=True):
    import dualvgr_videoqa_b200.model.models as M
    B, N, L, A, V, U = cfg
    model = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
    model.load_state_dict(orc.make_state_dict(U, A, V), strict=True)
    if p_zero:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
                m.dropout = 0.0
    return model.cuda().train(), [t.cuda() for t in orc.make_inputs(B, N, L, A, V)]


def test_fused_functions_match_the_module_by_module_path():
    """DualVGR.forward runs the question input unit and the unit stack as two fused autograd Functions; the mirrored modules
    keep their own forward() (one Function per module, reference call structure model/models.py:141-173). Same kernels
    underneath: outputs and every parameter gradient must agree to bf16 rounding of differently-ordered accumulations."""
    cfg = (5, 20, 9, 16, 50, 2)
    model, batch = _model(cfg)
    app, mot, q, qlen, ans = batch
    out = model(app, mot, q, qlen)
    loss = torch.nn.functional.cross_entropy(out[0], ans) + sum(t.float().pow(2).mean() for lst in out[3:] for t in lst)
    names = [n for n, _ in model.named_parameters()]
    g_fused = torch.autograd.grad(loss, list(model.parameters()), allow_unused=True)
    # module-by-module
    from dualvgr_videoqa_b200 import autograd as ag
    ag.begin_forward()
    q_emb, words, dq = model.linguistic_input_unit(q, qlen)
    a = model.visual_appearance_input_unit(app)
    B, N = mot.shape[:2]
    m_in = ag.ops.prep_features(mot.contiguous().view(B * N, -1), 1, False, False)
    m = ag.linear(m_in, model.visual_motion_input_unit.weight, model.visual_motion_input_unit.bias).view(B, N, -1)
    visual, aq, mq, ca, cm, aqf, mqf = model.visual_input_unit(a, m, dq, words, qlen)
    logits = model.output_unit(q_emb, model.feature_aggregation(visual))
    assert rel(out[0], logits) < 5e-3 and rel(out[1], aq) < 5e-3 and rel(out[2], mq) < 5e-3
    for x, y in zip(out[3] + out[4] + out[5] + out[6], ca + cm + aqf + mqf):
        assert rel(x, y) < 5e-3
    loss2 = torch.nn.functional.cross_entropy(logits, ans) + sum(t.float().pow(2).mean() for lst in (ca, cm, aqf, mqf) for t in lst)
    g_mod = torch.autograd.grad(loss2, list(model.parameters()), allow_unused=True)
    num = den = 0.0
    bad = []
    for n, g1, g2 in zip(names, g_fused, g_mod):
        assert (g1 is None) == (g2 is None), n
        if g1 is None:
            continue
        num += float((g1.double() - g2.double()).pow(2).sum()); den += float(g2.double().pow(2).sum())
        if float(g2.norm()) > 1e-3 and rel(g1, g2) > 5e-2:
            bad.append((n, round(rel(g1, g2), 4), float(g1.norm()), float(g2.norm())))
    assert not bad, bad
    assert (num / den) ** 0.5 < 1e-2, (num / den) ** 0.5


def test_lstm_dynamic_tile_schedule_survives_concurrent_launches(ops):
    """Two whole-sequence LSTM launches on two streams at once, each wanting every SM, plus a kernel that squats on most SMs
    while they start: with the dynamic tile claim no CTA ever waits on a tile owned by a CTA that is not running, so both
    finish with no dependency timeout and bit-identical results to a solo run."""
    g = torch.Generator().manual_seed(13)
    T, S, H, K1 = 6, 8192, 128, 256
    x = (torch.randn((T, S, K1), generator=g) * 0.5).to(BF16).cuda()
    wih = (torch.randn((2 * 4 * H, K1), generator=g) * 0.05).to(BF16).cuda()
    whh = (torch.randn((2, 4 * H, H), generator=g) * 0.05).to(BF16).cuda()
    bias = torch.zeros(2 * 4 * H, device="cuda")
    solo = ops.lstm_seq_fwd(x, wih, whh, bias)
    torch.cuda.synchronize()
    s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    big = torch.randn((64, 1 << 20), device="cuda")
    outs = []
    for rep in range(3):
        with torch.cuda.stream(s3):
            for _ in range(4):
                big = torch.sin(big) * 1.0001           # long elementwise kernels occupying SMs while the LSTMs launch
        with torch.cuda.stream(s1):
            a = ops.lstm_seq_fwd(x, wih, whh, bias)
        with torch.cuda.stream(s2):
            b = ops.lstm_seq_fwd(x, wih, whh, bias)
        outs += [a, b]
    torch.cuda.synchronize()
    for r in outs:
        assert int(r[5][-1]) == 0, "dependency poll timed out"
        assert torch.equal(r[3], solo[3]) and torch.equal(r[1][:, 1:], solo[1][:, 1:])      # (slot 0 of h_hist is never written)


def test_engine_state_dict_round_trip_and_shadow_sync():
    from dualvgr_videoqa_b200.engine import TrainEngine
    cfg = (4, 8, 6, 10, 30, 1)
    model, batch = _model(cfg)
    eng = TrainEngine(model, lr=1e-4)
    for _ in range(2):
        eng.train_step(*batch)
    sd_model = {k: v.clone() for k, v in model.state_dict().items()}
    sd_eng = eng.state_dict()
    l_next = float(eng.train_step(*batch))
    eng.close()
    model2, _ = _model(cfg)
    eng2 = TrainEngine(model2, lr=1e-4)
    model2.load_state_dict(sd_model)
    eng2.load_state_dict(sd_eng)                      # also re-casts the bf16 operand shadow from the restored weights
    assert torch.equal(eng2.shadow.float(), eng2.flat.to(BF16).float())
    l_resumed = float(eng2.train_step(*batch))
    assert abs(l_resumed - l_next) < 2e-3 * abs(l_next), (l_resumed, l_next)
    # a stale shadow is what sync_shadow() exists for: an external weight edit must reach the GEMM operands
    with torch.no_grad():
        model2.visual_motion_input_unit.weight.mul_(0.5)
    eng2.sync_shadow()
    assert torch.equal(eng2.shadow.float(), eng2.flat.to(BF16).float())
    eng2.close()


def test_dynamically_scheduled_gemm_equals_the_static_deal(ops):
    """dvgr_gemm with a tile counter (CTAs claim tiles) produces the same tiles as the static round-robin deal: bit-identical
    without split-K, fp32-atomic-order-identical within rounding with it; also while another stream squats on the SMs."""
    g = torch.Generator().manual_seed(17)
    M, N, K = 3000, 1536, 2048
    a = (torch.randn((M, K), generator=g) * 0.1).to(BF16).cuda()
    w = (torch.randn((N, K), generator=g) * 0.05).to(BF16).cuda()
    c0 = torch.empty((M, N), dtype=BF16, device="cuda")
    c1 = torch.empty_like(c0)
    ops.gemm(a, 0, w, 0, M, N, K, c0)
    side = torch.cuda.Stream()
    big = torch.randn((64, 1 << 20), device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(3):
            big = torch.sin(big) * 1.0001
    ops.gemm(a, 0, w, 0, M, N, K, c1, dynamic=True)
    torch.cuda.synchronize()
    assert torch.equal(c0, c1)
    # weight-gradient form: MN-major operands, split-K with fp32 atomics into a zeroed buffer
    dy = (torch.randn((20000, 768), generator=g) * 0.1).to(BF16).cuda()
    x = (torch.randn((20000, 512), generator=g) * 0.1).to(BF16).cuda()
    o0, o1 = torch.zeros((768, 512), device="cuda"), torch.zeros((768, 512), device="cuda")
    ops.linear_wgrad(dy, x, out=o0, atomic=True)
    ops.linear_wgrad(dy, x, out=o1, atomic=True, dynamic=True)
    ref = dy.double().t() @ x.double()
    assert rel(o0, ref) < 1e-5 and rel(o1, ref) < 1e-5
