"""Per-kernel parity: every fused CUDA kernel (through the C ABI) against the CPU oracle's formula for the same op,
forward and backward (oracle gradients via autograd in float64 on the SAME bf16-rounded inputs).

Tolerance: bf16 storage of outputs => relative L2 error <= 1e-2 (north_star: 2e-2 in bf16); fp32-output kernels 1e-4."""
import math

import pytest
import torch

import dualvgr_oracle as orc

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf(x):
    """round to bf16 and return (cuda bf16 tensor, float64 cpu copy with grad)."""
    xb = x.to(BF16)
    return xb.cuda(), xb.double().requires_grad_(True)


@pytest.fixture(scope="module")
def ops():
    import dualvgr_videoqa_b200.ops as ops
    import dualvgr_videoqa_b200._lib as L
    L.lib.dvgr_set_seed_offset(None)        # a train engine of an earlier test may have installed its device counter
    return ops


@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (200, 136, 72), (1000, 32, 1536)])
def test_gemm_forward_dgrad_wgrad(ops, M, N, K):
    torch.manual_seed(0)
    x, xr = bf(torch.randn(M, K))
    w, wr = bf(torch.randn(N, K) * 0.05)
    b = torch.randn(N)
    y = ops.linear_fwd(x, w, bias=b.cuda(), act="elu")
    ref = torch.nn.functional.elu(xr @ wr.t() + b.double())
    assert rel(y, ref) < 1e-2
    dy, dyr = bf(torch.randn(M, N))
    assert rel(ops.linear_dgrad(dy, w), dyr.detach() @ wr.detach()) < 1e-2
    assert rel(ops.linear_wgrad(dy, x), dyr.detach().t() @ xr.detach()) < 1e-4
    # device-side SIMT reference agrees too (used for the full-size checks)
    c = ops.gemm_reference(x, K, 1, w, K, 1, M, N, K)
    assert rel(c, xr.detach() @ wr.detach().t()) < 1e-5


def test_grouped_weight_gradients(ops):
    """dvgr_wgrad_grouped: several independent dW += dy^T x problems of different (ragged) shapes in ONE persistent launch,
    against float64 and against the one-launch-per-problem path; accumulation into non-zero buffers; > 32 problems (two launches)."""
    torch.manual_seed(5)
    shapes = [(5120, 768, 768), (1000, 136, 72), (256, 1536, 300), (700, 4002, 768), (10240, 768, 768), (64, 8, 8)]
    shapes = shapes + [(300 + 17 * i, 128, 64) for i in range(30)]
    refs, outs, keep = [], [], []
    for (M, N, K) in shapes:
        N8, K8 = (N + 7) // 8 * 8, (K + 7) // 8 * 8
        dy = torch.zeros(M, N8); dy[:, :N] = torch.randn(M, N) * 0.5
        x = torch.zeros(M, K8); x[:, :K] = torch.randn(M, K) * 0.5
        dyb, xb = dy.to(BF16).cuda(), x.to(BF16).cuda()
        base = torch.randn(N, K).cuda()
        out = base.clone()
        ops.wgrad_enqueue(dyb, xb, out, rows=N, cols=K)
        refs.append(base.double().cpu() + dyb.double().cpu()[:, :N].t() @ xb.double().cpu()[:, :K])
        outs.append(out); keep.append((dyb, xb, base))
    assert ops.flush_wgrads() == len(shapes)
    torch.cuda.synchronize()
    for o, r, (M, N, K) in zip(outs, refs, shapes):
        assert rel(o, r) < 1e-4, (M, N, K)
    # the single-problem split-K GEMM computes the same products
    for (dyb, xb, base), o, (M, N, K) in list(zip(keep, outs, shapes))[:4]:
        single = base.clone()
        ops.linear_wgrad(dyb, xb, out=single, atomic=True, rows=N, cols=K)
        assert rel(single, o) < 1e-5


def test_grouped_column_sums(ops):
    """dvgr_colsum_grouped: the queued bias gradients of a step in one launch — bf16 and fp32 inputs, strided views, ragged
    widths, accumulation into non-zero outputs, more problems than one launch holds."""
    torch.manual_seed(6)
    shapes = [(5120, 768, BF16), (10240, 768, BF16), (256, 1536, BF16), (300, 4008, BF16), (77, 13, torch.float32), (1, 8, BF16)]
    shapes += [(100 + 31 * i, 64 + 8 * (i % 5), BF16) for i in range(50)]
    outs, refs = [], []
    for (R, C, dt) in shapes:
        full = (torch.randn(R, C + 8) * 0.5).to(dt).cuda()
        x = full[:, :C]                                   # strided view (row stride C + 8)
        base = torch.randn(C).cuda()
        out = base.clone()
        ops.colsum_enqueue(x, out)
        outs.append(out); refs.append(base.double().cpu() + x.double().cpu().sum(0))
    ops.flush_wgrads()
    torch.cuda.synchronize()
    for o, r, sh in zip(outs, refs, shapes):
        assert rel(o, r) < 1e-5, sh


def test_gemm_full_size_property(ops):
    """BASELINE config-2 shape of the appearance W_ih product; linearity property + sampled rows vs the SIMT reference."""
    torch.manual_seed(1)
    M, N, K = 81920, 3072, 2048
    x = (torch.randn(M, K, device="cuda") * 0.5).to(BF16)
    w = (torch.randn(N, K, device="cuda") * 0.02).to(BF16)
    y = ops.linear_fwd(x, w, out_dtype=torch.float32)
    idx = torch.randint(0, M, (256,), device="cuda")
    ref = ops.gemm_reference(x[idx].contiguous(), K, 1, w, K, 1, 256, N, K)
    assert rel(y[idx], ref) < 1e-5
    # linearity: (2x) W = 2 (x W) exactly in fp32 accumulation of bf16 products (power-of-two scaling)
    y2 = ops.linear_fwd((x * 2).to(BF16), w, out_dtype=torch.float32)
    assert torch.equal(y2[idx], 2 * y[idx])


@pytest.mark.parametrize("T,S,H,D", [(4, 200, 128, 2), (16, 160, 384, 2), (3, 130, 64, 4), (16, 2560, 384, 2)])
def test_lstm_fused_recurrence(ops, T, S, H, D):
    torch.manual_seed(2)
    gx, gxr = bf(torch.randn(T, S, D * 4 * H))
    whh, whhr = bf(torch.randn(D, 4 * H, H) * 0.08)

    def ref_run(gx, whh):
        hs = []
        for d in range(D):
            h = gx.new_zeros(S, H); c = gx.new_zeros(S, H)
            for s in range(T):
                t = s if d % 2 == 0 else T - 1 - s
                pre = (gx[t, :, d * 4 * H:(d + 1) * 4 * H] + h @ whh[d].t()).view(S, H, 4)
                i, f, g, o = pre[..., 0].sigmoid(), pre[..., 1].sigmoid(), pre[..., 2].tanh(), pre[..., 3].sigmoid()
                c = f * c + i * g
                h = o * c.tanh()
            hs.append(h)
        return torch.cat(hs, 1)

    ref = ref_run(gxr, whhr)
    g = gx.clone()
    h_hist, c_hist, h_last, _ = ops.lstm_fwd(g, whh)
    assert rel(h_last, ref) < 1e-2
    dh, dhr = bf(torch.randn(S, D * H))
    ref.backward(dhr.detach())
    g_act = g.clone()
    ops.lstm_bwd(g, whh, h_hist, c_hist, dh)
    assert rel(g, gxr.grad) < 2e-2
    # whole-sequence persistent backward (ONE launch for steps T-2..0, cross-CTA step chaining) = the per-step path
    # (it consumes the kernels' blocked layout of the activated gates / cell states: re-layout the per-step tensors)
    gb, cb = ops.lstm_block_gates(g_act, D), ops.lstm_block_c(c_hist)
    assert torch.equal(ops.lstm_unblock_gates(gb, S), g_act) and torch.equal(ops.lstm_unblock_c(cb, S), c_hist)
    g2, sync = ops.lstm_bwd(gb, whh, h_hist, cb, dh, whole_sequence=True)
    torch.cuda.synchronize()
    assert int(sync[-1]) == 0, "dependency poll timed out"
    # sync = per-(direction, block) completion counters | tile-claim counter | sticky error word
    assert int(sync[:-2].min()) == int(sync[:-2].max()) == 16 * ((H + 127) // 128) * (T - 1)
    assert rel(g2, gxr.grad) < 2e-2
    assert torch.equal(g2, g)          # same arithmetic in the same order: bit-identical to the step launches
    g3, _ = ops.lstm_bwd(gb, whh, h_hist, cb, dh, whole_sequence=True)
    assert torch.equal(g3, g2)


@pytest.mark.parametrize("T,S,H,D,K1", [(4, 200, 128, 2, 72), (16, 160, 384, 2, 2048), (5, 300, 64, 4, 304),
                                        (16, 2560, 384, 2, 256)])
def test_lstm_whole_sequence_fused_forward(ops, T, S, H, D, K1):
    """dvgr_lstm_seq_fwd (input projection + all steps in ONE persistent launch, cross-CTA step dependencies) against the
    float64 recurrence and against the per-step path; the last shape has more tiles per step than SMs (multi-wave
    dependency chains); the sticky dependency-timeout word must stay 0."""
    torch.manual_seed(21)
    x, xr = bf(torch.randn(T, S, K1) * 0.5)
    wih, wihr = bf(torch.randn(D * 4 * H, K1) * (0.5 / math.sqrt(K1)))
    whh, whhr = bf(torch.randn(D, 4 * H, H) * 0.08)
    bias = torch.randn(D * 4 * H) * 0.1
    gxr = (xr.detach() @ wihr.detach().t() + bias.double())           # [T, S, D*4H], interleaved gate columns

    def ref_run():
        hs, acts = [], torch.zeros(T, S, D * 4 * H, dtype=torch.float64)
        for d in range(D):
            h = gxr.new_zeros(S, H); c = gxr.new_zeros(S, H)
            for s in range(T):
                t = s if d % 2 == 0 else T - 1 - s
                pre = (gxr[t, :, d * 4 * H:(d + 1) * 4 * H] + h @ whhr[d].detach().t()).view(S, H, 4)
                i, f, g, o = pre[..., 0].sigmoid(), pre[..., 1].sigmoid(), pre[..., 2].tanh(), pre[..., 3].sigmoid()
                acts[t, :, d * 4 * H:(d + 1) * 4 * H] = torch.stack([i, f, g, o], -1).view(S, 4 * H)
                c = f * c + i * g
                h = o * c.tanh()
            hs.append(h)
        return torch.cat(hs, 1), acts

    ref_h, ref_g = ref_run()
    gates, h_hist, c_hist, h_last, _, sync = ops.lstm_seq_fwd(x, wih, whh, bias.cuda())
    torch.cuda.synchronize()
    assert int(sync[-1]) == 0, "dependency poll timed out"
    assert int(sync[:-2].min()) == int(sync[:-2].max()) == 16 * (4 * H // 256) * T     # every (warp, tile) published once
    n_tiles = D * ((S + 127) // 128) * (4 * H // 256) * T
    assert int(sync[-2]) >= n_tiles                                                    # every tile was claimed (+ one miss per CTA)
    assert rel(h_last, ref_h) < 1e-2
    assert rel(ops.lstm_unblock_gates(gates, S), ref_g) < 1e-2          # activated gates, stored in the blocked layout
    # the per-step path on the same operands agrees (it rounds the pre-activations to bf16 first, so not bit-equal)
    g2 = ops.linear_fwd(x.view(T * S, K1), wih, bias=bias.cuda()).view(T, S, D * 4 * H)
    _, c2, h2, _ = ops.lstm_fwd(g2, whh)
    # (slot 0 = the initial state is implicit zeros in the whole-sequence layout and never written: compare slots 1..T)
    assert rel(h_last, h2) < 1e-2 and rel(ops.lstm_unblock_c(c_hist, S)[:, 1:], c2[:, 1:]) < 1e-2
    # idempotence: a second launch on fresh sync words reproduces the result bit for bit (no race on the step chain)
    gates_b, _, c_b, h_b, _, _ = ops.lstm_seq_fwd(x, wih, whh, bias.cuda())
    assert torch.equal(h_b, h_last)
    assert torch.equal(ops.lstm_unblock_gates(gates_b, S), ops.lstm_unblock_gates(gates, S))
    assert torch.equal(ops.lstm_unblock_c(c_b, S)[:, 1:], ops.lstm_unblock_c(c_hist, S)[:, 1:])


@pytest.mark.parametrize("B,N,p", [(3, 20, 0.0), (5, 8, 0.0), (2, 33, 0.0), (2, 64, 0.0)])
def test_gat_attention_fwd_bwd(ops, B, N, p):
    torch.manual_seed(3)
    D, K = 768, 4
    Dh = D // K
    sd = {}
    wh_list, whr_list, av_list = [], [], []
    adj = orc.build_adjacency(N)
    if N == 33:
        adj[3] = 0          # fully masked row -> uniform attention
        adj[5, ::2] = 0
    gate, gater = torch.rand(B, N), None
    gater = gate.double().requires_grad_(True)
    outs_ref, avr_list = [], []
    for g in range(2):
        wh, whr = bf(torch.randn(B * N, D))
        av = torch.randn(K, 2 * Dh + 1) * 0.1
        avr = av.double().requires_grad_(True)
        wh_list.append(wh); whr_list.append(whr); av_list.append(av.cuda()); avr_list.append(avr)
        heads = []
        W = whr.view(B, N, K, Dh)
        for k in range(K):
            Whk = W[:, :, k]
            s = Whk @ avr[k, :Dh]
            t = Whk @ avr[k, Dh:2 * Dh]
            e = torch.nn.functional.leaky_relu(s[:, :, None] + t[:, None, :] + avr[k, 2 * Dh], 0.01)
            e = torch.where(adj > 0, e, torch.full_like(e, -9e15))
            P = torch.softmax(e, -1)
            heads.append(torch.nn.functional.elu(P @ (Whk * gater[:, :, None])))
        outs_ref.append(torch.cat(heads, -1))
    gc = gate.cuda()
    outs, outs32 = ops.gat_attn_fwd(wh_list, [gc, gc], av_list, adj.cuda(), B, N, want_f32=True)
    for g in range(2):
        assert rel(outs[g].view(B, N, D), outs_ref[g]) < 1e-2
        assert rel(outs32[g], outs_ref[g]) < 1e-4
    douts, doutr = [], []
    for g in range(2):
        d, dr = bf(torch.randn(B * N, D))
        douts.append(d); doutr.append(dr.detach().view(B, N, D))
    d32 = [torch.randn(B, N, D) * 0.5 for _ in range(2)]
    loss = sum((outs_ref[g] * (doutr[g] + d32[g].double())).sum() for g in range(2))
    loss.backward()
    dwhs, dgates, davecs = ops.gat_attn_bwd(wh_list, [gc, gc], av_list, outs, douts, adj.cuda(), B, N,
                                             douts32=[t.cuda() for t in d32])
    for g in range(2):
        assert rel(dwhs[g], whr_list[g].grad) < 2e-2
        assert rel(davecs[g], avr_list[g].grad) < 2e-2
    assert rel(dgates[0] + dgates[1], gater.grad) < 2e-2


@pytest.mark.parametrize("B,N", [(4, 20), (3, 40), (2, 8)])
def test_gat_tensor_core_path_matches_generic_path_with_dropout(ops, B, N):
    """The mma.sync fast path (D=768, 4 heads) and the generic SIMT kernels draw the SAME dropout masks (attention and
    output) from (seed, stream, element): with p > 0 the two paths must agree element-wise, forward and backward, and
    drop exactly the same outputs. N=40 exercises the 64-node forward variant (its backward stays generic)."""
    import os
    torch.manual_seed(31)
    D, K = 768, 4
    Dh = D // K
    adj = orc.build_adjacency(N).cuda()
    gate = torch.rand(B, N).cuda()
    whs = [(torch.randn(B * N, D)).to(BF16).cuda() for _ in range(2)]
    avs = [(torch.randn(K, 2 * Dh + 1) * 0.1).cuda() for _ in range(2)]
    douts = [(torch.randn(B * N, D)).to(BF16).cuda() for _ in range(2)]
    d32 = [(torch.randn(B, N, D) * 0.5).cuda() for _ in range(2)]
    res = {}
    try:
        for mode in ("0", "3"):
            os.environ["DVGR_GAT_FAST"] = mode
            outs, o32 = ops.gat_attn_fwd(whs, [gate, gate], avs, adj, B, N, p_att=0.3, p_out=0.3, seed=77, want_f32=True)
            dwhs, dgs, dav = ops.gat_attn_bwd(whs, [gate, gate], avs, outs, douts, adj, B, N, p_att=0.3, p_out=0.3,
                                              seed=77, douts32=d32)
            torch.cuda.synchronize()
            res[mode] = (outs, o32, dwhs, dgs, dav)
    finally:
        os.environ.pop("DVGR_GAT_FAST", None)
    for g in range(2):
        a, b_ = res["0"], res["3"]
        assert torch.equal(a[1][g] == 0, b_[1][g] == 0)                       # identical output-dropout pattern
        drop = float((b_[1][g] == 0).float().mean())
        assert abs(drop - 0.3) < 0.03
        assert rel(b_[1][g], a[1][g]) < 1e-4 and rel(b_[0][g], a[0][g]) < 1e-2
        assert rel(b_[2][g], a[2][g]) < 2e-2 and rel(b_[3][g], a[3][g]) < 2e-2 and rel(b_[4][g], a[4][g]) < 2e-2


def test_gat_dropout_statistics(ops):
    """Train-mode dropout is a Philox stream of our own: check keep-rate and scaling, and fwd/bwd mask consistency."""
    torch.manual_seed(4)
    B, N, D = 64, 20, 768
    x = torch.ones(B * N * D // 8 * 8).to(BF16).cuda()
    y = ops.dropout_raw(x, 0.15, seed=123, stream_id=7)
    keep = (y != 0).float().mean().item()
    assert abs(keep - 0.85) < 5e-3
    assert torch.allclose(y[y != 0].float(), torch.tensor(1 / 0.85).to(BF16).float().cuda())
    y2 = ops.dropout_raw(x, 0.15, seed=123, stream_id=7)
    assert torch.equal(y, y2)                                   # same (seed, stream) -> same mask (backward relies on it)
    y3 = ops.dropout_raw(x, 0.15, seed=124, stream_id=7)
    assert not torch.equal(y, y3)


def test_qattn_and_gates(ops):
    torch.manual_seed(5)
    B, L, D, W, N = 6, 9, 768, 300, 20
    y, yr = bf(torch.randn(B, L, D))
    yr.data[1, 5:] = 0                                           # padded rows of dynamic_q give y = bias only; keep generic
    y = yr.detach().to(BF16).cuda()
    words_full = torch.zeros(B, L, 304)
    words_full[:, :, :W] = torch.randn(B, L, W).tanh()
    words, wordsr = bf(words_full)
    wf = torch.randn(D) * 0.05
    cf = torch.randn(1)
    qlen = torch.tensor([9, 5, 3, 7, 9, 4])
    wfr, cfr = wf.double().requires_grad_(True), cf.double().requires_grad_(True)
    d = yr / yr.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    score = d @ wfr + cfr
    prob = torch.softmax(score, 1)
    mask = (torch.arange(L)[None] < qlen[:, None]).double()
    al = prob * mask
    al = al / (al.sum(1, keepdim=True) + 1e-5)
    qc_ref = torch.einsum("bl,blw->bw", al, wordsr)
    qc, alpha, nrm, prob_k, ssum = ops.qattn_fwd(y, wf.cuda(), cf.cuda(), qlen.int().cuda(), words, W, 304)
    assert rel(alpha, al) < 1e-4
    assert rel(qc[:, :W], qc_ref[:, :W]) < 1e-2
    assert float(qc[:, W:].abs().max()) == 0.0
    dqc, dqcr = bf(torch.randn(B, 304))
    (qc_ref * dqcr.detach()).sum().backward()
    dy, dwords, dwf, dcf = ops.qattn_bwd(dqc, y, wf.cuda(), qlen.int().cuda(), words, W, alpha, nrm, prob_k, ssum)
    m = yr.detach().norm(dim=-1) > 0
    assert rel(dy[m.cuda()], yr.grad[m]) < 2e-2
    assert rel(dwords[:, :, :W], wordsr.grad[:, :, :W]) < 2e-2
    assert rel(dwf, wfr.grad) < 2e-2
    # gates
    xa, xar = bf(torch.randn(B, N, D) * 0.3)
    xm, xmr = bf(torch.randn(B, N, D) * 0.3)
    q, qr = bf(torch.randn(B, 2 * D) * 0.1)
    ga_ref = torch.sigmoid(torch.einsum("bnd,bd->bn", xar, qr[:, :D]))
    gm_ref = torch.sigmoid(torch.einsum("bnd,bd->bn", xmr, qr[:, D:]))
    ga, gm = ops.gate_fwd(xa, xm, q)
    assert rel(ga, ga_ref) < 1e-3 and rel(gm, gm_ref) < 1e-3
    dga, dga2, dgm = torch.randn(B, N), torch.randn(B, N), torch.randn(B, N)
    ((ga_ref * (dga + dga2).double()).sum() + (gm_ref * dgm.double()).sum()).backward()
    dxa, dxm = torch.zeros_like(xa), torch.zeros_like(xm)
    dq = ops.gate_bwd(xa, xm, q, ga, gm, dga.cuda(), dga2.cuda(), dgm.cuda(), None, dxa, dxm)
    assert rel(dxa, xar.grad) < 2e-2 and rel(dxm, xmr.grad) < 2e-2
    assert rel(dq, qr.grad) < 2e-2


def test_view_attention_mfb_readout(ops):
    torch.manual_seed(6)
    B, N, D = 4, 20, 768
    M = B * N
    hid, hidr = bf(torch.randn(2, M, D).tanh())
    z, zr = bf(torch.randn(2, M, D))
    x, xr = bf(torch.randn(M, D))
    w2 = torch.randn(D) * 0.05
    w2r = w2.double().requires_grad_(True)
    wv = hidr @ w2r
    beta = torch.softmax(wv, 0)
    emb_ref = (beta.unsqueeze(-1) * zr).sum(0)
    xnew_ref = xr + emb_ref
    xnew, emb, beta_k = ops.view_attn_fwd(hid, z, x, w2.cuda())
    assert rel(xnew, xnew_ref) < 1e-2 and rel(emb, emb_ref) < 1e-2
    dxn, dxnr = bf(torch.randn(M, D))
    de, der = bf(torch.randn(M, D))
    ((xnew_ref * dxnr.detach()).sum() + (emb_ref * der.detach()).sum()).backward()
    dz, dhid, dw2 = ops.view_attn_bwd(dxn, de, hid, z, w2.cuda(), beta_k)
    assert rel(dz, zr.grad) < 2e-2
    # hidden = tanh(pre): the kernel returns d pre
    assert rel(dhid, hidr.grad * (1 - hidr.detach() ** 2)) < 2e-2
    assert rel(dw2, w2r.grad) < 2e-2

    # MFB pair-sum
    x0, x0r = bf(torch.nn.functional.elu(torch.randn(M, 512)))
    x1, x1r = bf(torch.nn.functional.elu(torch.randn(M, 512)))
    zz_ref = (x0r * x1r).view(M, 256, 2).sum(-1)
    zz = ops.mfb_fwd(x0, x1)
    assert rel(zz, zz_ref) < 1e-2
    dzz, dzzr = bf(torch.randn(M, 256))
    (zz_ref * dzzr.detach()).sum().backward()
    d0, d1 = ops.mfb_bwd(dzz, x0, x1)
    eg = lambda y: torch.where(y > 0, torch.ones_like(y), y + 1)
    assert rel(d0, x0r.grad * eg(x0r.detach())) < 2e-2 and rel(d1, x1r.grad * eg(x1r.detach())) < 2e-2

    # read-out
    v, vr = bf(torch.randn(B, N, D))
    u, ur = bf(torch.nn.functional.elu(torch.randn(B, N, D)))
    w = torch.randn(D) * 0.05
    c = torch.randn(1)
    wr = w.double().requires_grad_(True)
    al = torch.softmax(ur @ wr + c.double(), 1)
    pooled_ref = (al.unsqueeze(-1) * vr).sum(1)
    pooled, alpha = ops.readout_fwd(v, u, w.cuda(), c.cuda())
    assert rel(pooled, pooled_ref) < 1e-2
    dp, dpr = bf(torch.randn(B, D))
    (pooled_ref * dpr.detach()).sum().backward()
    dv, du, dw, dc = ops.readout_bwd(dp, v, u, w.cuda(), alpha)
    assert rel(dv, vr.grad) < 2e-2
    assert rel(du, ur.grad * eg(ur.detach())) < 2e-2
    assert rel(dw, wr.grad) < 2e-2


def test_bn_ce_colsum(ops):
    torch.manual_seed(7)
    B, D, A = 37, 768, 4002
    x, xr = bf(torch.randn(B, D) * 2 + 0.5)
    gamma, betap = torch.rand(D) + 0.5, torch.randn(D) * 0.1
    gr, br = gamma.double().requires_grad_(True), betap.double().requires_grad_(True)
    rm, rv = torch.zeros(D), torch.ones(D)
    mean, var = xr.mean(0), xr.var(0, unbiased=False)
    y_ref = (xr - mean) / torch.sqrt(var + 1e-5) * gr + br
    rmc, rvc = rm.cuda(), rv.cuda()
    y, m, rs = ops.bn_fwd(x, gamma.cuda(), betap.cuda(), rmc, rvc, True)
    assert rel(y, y_ref) < 1e-2
    assert rel(rmc, 0.1 * mean.detach()) < 1e-3
    assert rel(rvc, 0.9 + 0.1 * xr.detach().var(0, unbiased=True)) < 1e-3
    dy, dyr = bf(torch.randn(B, D))
    (y_ref * dyr.detach()).sum().backward()
    dx, dg, db = ops.bn_bwd(dy, x, gamma.cuda(), m, rs, True)
    assert rel(dx, xr.grad) < 2e-2 and rel(dg, gr.grad) < 1e-3 and rel(db, br.grad) < 1e-3
    # eval mode uses running stats
    ye, _, _ = ops.bn_fwd(x, gamma.cuda(), betap.cuda(), rmc, rvc, False)
    ye_ref = (xr.detach() - rmc.cpu().double()) / torch.sqrt(rvc.cpu().double() + 1e-5) * gamma.double() + betap.double()
    assert rel(ye, ye_ref) < 1e-2
    # cross entropy
    logits = torch.randn(B, A)
    lr = logits.double().requires_grad_(True)
    ans = torch.randint(0, A, (B,))
    ce_ref = torch.nn.functional.cross_entropy(lr, ans)
    ce_ref.backward()
    loss, dlog, correct = ops.cross_entropy(logits.cuda(), ans.cuda())
    assert abs(float(loss) - float(ce_ref)) < 1e-5 * abs(float(ce_ref))
    assert rel(dlog[:, :A], lr.grad) < 1e-2
    assert float(dlog[:, A:].abs().max()) == 0.0
    assert torch.equal(correct.cpu().bool(), logits.argmax(1) == ans)
    # colsum
    t = torch.randn(1000, 300)
    assert rel(ops.colsum(t.cuda()), t.double().sum(0)) < 1e-5
    tb = t.to(BF16)
    assert rel(ops.colsum(tb.cuda()), tb.double().sum(0)) < 1e-5


@pytest.mark.parametrize("N", [8, 20])
def test_pair_losses(ops, N):
    torch.manual_seed(8)
    B, D = 5, 768
    # node-diverse AND the ill-conditioned regime of the real model (nearly identical nodes)
    for spread in (1.0, 1e-3):
        base = torch.randn(B, 1, D)
        x = (base + spread * torch.randn(B, N, D)).float()
        y = (base * 0.7 + spread * torch.randn(B, N, D)).float()
        xr, yr = x.double().requires_grad_(True), y.double().requires_grad_(True)
        com = orc.common_loss(xr, yr)
        gx, gy = torch.autograd.grad(com, (xr, yr))
        loss, dx, dy = ops.pair_loss(x.cuda(), y.cuda(), 0, 1.0 / (B * N * N))
        # N x N Grams and gradient products run on TF32 tensor cores (10-bit mantissa): 1e-3; the ill-conditioned regime
        # (nearly identical nodes) sits at the fp32 noise floor of the centred/normalised form
        tol = 1e-3 if spread == 1.0 else 5e-2
        assert abs(float(loss) - float(com)) <= tol * abs(float(com))
        assert rel(dx, gx) < max(tol, 1e-3) and rel(dy, gy) < max(tol, 1e-3)
        hs = orc.loss_dependence(xr, yr, N)
        gx, gy = torch.autograd.grad(hs, (xr, yr))
        loss, dx, dy = ops.pair_loss(x.cuda(), y.cuda(), 1, 1.0)
        assert abs(float(loss) - float(hs)) <= tol * abs(float(hs))
        assert rel(dx, gx) < max(tol, 1e-3) and rel(dy, gy) < max(tol, 1e-3)


@pytest.mark.parametrize("N", [8, 20, 33])
def test_unit_aux_losses_tensor_centric(ops, N):
    """dvgr_aux_loss_unit: common(ca, cm) + HSIC(aq, ca) + HSIC(mq, cm) of one DualVGR unit, value and the four complete
    gradients, against the oracle's formulas (float64 autograd) and against the pair-centric kernels."""
    torch.manual_seed(18)
    B, D = 4, 768
    c_com, c_dep = 1.0 / (B * N * N), 3e-3
    for spread in (1.0, 1e-3):
        base = torch.randn(B, 1, D)
        ts = [(base * s + spread * torch.randn(B, N, D)).float() for s in (1.0, 0.7, 0.9, 1.2)]
        rs = [t.double().requires_grad_(True) for t in ts]
        com = orc.common_loss(rs[0], rs[1]) * (c_com * B * N * N)
        h1, h2 = orc.loss_dependence(rs[2], rs[0], N) * c_dep, orc.loss_dependence(rs[3], rs[1], N) * c_dep
        grads = torch.autograd.grad(com + h1 + h2, rs)
        vals, got = ops.aux_loss_unit(*[t.cuda() for t in ts], c_com, c_dep)
        tol = 1e-3 if spread == 1.0 else 5e-2
        for v, r in zip(vals.cpu(), (com, h1, h2)):
            assert abs(float(v) - float(r)) <= tol * abs(float(r))
        for gg, gr in zip(got, grads):
            assert rel(gg, gr) < max(tol, 1e-3)
        # the pair-centric kernels (still behind utils.common_loss / loss_dependence) agree
        _, dx, dy = ops.pair_loss(ts[0].cuda(), ts[1].cuda(), 0, c_com)
        _, dxa, dya = ops.pair_loss(ts[2].cuda(), ts[0].cuda(), 1, c_dep)
        assert rel(got[0], dx + dya) < max(tol, 2e-3) and rel(got[2], dxa) < max(tol, 2e-3)


def test_prep_cast_adam(ops):
    torch.manual_seed(9)
    S, T, C = 12, 16, 2048
    x = torch.randn(S * T, C).abs()
    out = ops.prep_features(x.cuda(), T, True, True)
    ref = x.tanh().view(S, T, C).transpose(0, 1).reshape(T * S, C)
    assert rel(out, ref) < 5e-3
    w = torch.randn(4 * 384, 300)
    wc = ops.cast_rows(w.cuda(), out_cols=304, lstm_H=384)
    ref = w.view(4, 384, 300).permute(1, 0, 2).reshape(4 * 384, 300)
    assert rel(wc[:, :300], ref) < 5e-3 and float(wc[:, 300:].abs().max()) == 0.0
    # fused clip + Adam vs torch
    n = 100003
    p0, g0 = torch.randn(n), torch.randn(n) * 3
    pr = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pr], lr=1e-4)
    p, m, v = p0.cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
    for step in range(1, 4):
        g = g0 * step
        pr.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([pr], max_norm=12)
        opt.step()
        gc = g.cuda()
        nsq = ops.sumsq(gc)
        assert abs(float(nsq) - float(g.double().pow(2).sum())) < 1e-4 * float(g.double().pow(2).sum())
        ops.adam_step(p, gc, m, v, 1e-4, step, max_norm=12.0, norm_sq=nsq)
    assert rel(p, pr.detach()) < 1e-6


def test_masked_four_direction_lstm_matches_oracle(ops):
    """Question-encoder recurrence: 4 directions, per-sequence lengths, per-step outputs, padded-step gradient carry.
    Checked against the oracle's step-by-step LSTM (which restates pack_padded_sequence semantics)."""
    import dualvgr_videoqa_b200.autograd as ag
    torch.manual_seed(11)
    B, L, W, H = 70, 7, 300, 384
    words = torch.randn(B, L, W).tanh()
    qlen = torch.randint(1, L + 1, (B,)); qlen[0] = L; qlen[1] = 1
    params = []
    for _ in range(2):
        for _d in range(2):
            params += [torch.randn(4 * H, W) * 0.05, torch.randn(4 * H, H) * 0.05, torch.randn(4 * H) * 0.1, torch.randn(4 * H) * 0.1]
    pb = [p.to(BF16).float() for p in params]                      # weights as the kernel sees them (bf16-rounded)
    for i in (2, 3, 6, 7, 10, 11, 14, 15):
        pb[i] = params[i]                                           # biases stay fp32
    wb = words.to(BF16).float()
    cp = [p.clone().cuda().requires_grad_(True) for p in pb]
    wc = wb.clone().cuda().requires_grad_(True)
    dq, q = ag.QuestionEncoderFn.apply(wc, qlen.int().cuda(), *cp)
    rp = [p.double().requires_grad_(True) for p in pb]
    wr = wb.double().requires_grad_(True)
    outs = []
    for base in (0, 8):
        of, hf = orc.lstm_direction(wr, rp[base], rp[base + 1], rp[base + 2], rp[base + 3], False, qlen)
        ob, hb = orc.lstm_direction(wr, rp[base + 4], rp[base + 5], rp[base + 6], rp[base + 7], True, qlen)
        outs.append((torch.cat([of, ob], -1), torch.cat([hf, hb], -1)))
    dq_ref, q_ref = outs[0][0], outs[1][1]
    assert rel(dq, dq_ref) < 1e-2 and rel(q, q_ref) < 1e-2
    pad = torch.arange(L)[None, :] >= qlen[:, None]
    assert float(dq.float().cpu()[pad].abs().max()) == 0.0
    g1, g1r = bf(torch.randn(B, L, 2 * H))
    g2, g2r = bf(torch.randn(B, 2 * H))
    ((dq_ref * g1r.detach()).sum() + (q_ref * g2r.detach()).sum()).backward()
    ((dq.float() * g1.float()).sum() + (q.float() * g2.float()).sum()).backward()
    assert rel(wc.grad, wr.grad) < 2e-2
    for i, (a, b) in enumerate(zip(cp, rp)):
        assert rel(a.grad, b.grad) < 2e-2, i
