"""fp32 mode of the hot path (north_star: "within 1e-4 relative error in fp32").

Every activation is fp32; the streaming kernels run in their fp32-activation build (csrc/act.cuh, entry points *_f32) and
every matrix product runs on the tcgen05 GEMM as a 3 x bf16 split product (ops.split3 + ops.gemm3: a_lo b_hi + a_hi b_hi +
a_hi b_lo accumulated in fp32 TMEM, < 2e-5 of the float64 product per GEMM). The recurrences run step by step: one batched
split product h_{t-1} W_hh^T and one fp32 cell launch (csrc/lstm32.cu) per time step.

The autograd Functions here are the fp32 counterparts of autograd.LinearFn / AppearanceEncoderFn / QuestionEncoderFn /
GatLayerFn; the dtype-agnostic Functions (QAttnFn, GateFn, ViewAttnFn, MfbPairFn, ReadoutFn, BatchNormFn, DropoutFn) serve
both precisions. autograd.linear / fused_gat_layer dispatch here on the activation dtype. fp32 mode favours accuracy over
speed: no stacked layouts, no deferred weight gradients, no side streams."""
import ctypes
import weakref

import torch
from torch.autograd import Function

from . import _lib, ops
from . import autograd as ag

BF16, F32 = torch.bfloat16, torch.float32

_pcache = {}


def weight_planes(params, tag=""):
    """bf16 planes [3, rows, Cp] (lo | hi | hi) of the row-wise concatenation of fp32 matrices; cached per parameter object
    and re-split when a version counter or the optimizer epoch moves (same policy as autograd.bf16_rows)."""
    key = (tag, tuple(id(p) for p in params))
    ver = ag._versions(params)
    ent = _pcache.get(key)
    if ent is not None and ent[0] == ver and all(r() is p for r, p in zip(ent[2], params)):
        return ent[1]
    with torch.no_grad():
        w = params[0].detach() if len(params) == 1 else torch.cat([p.detach() for p in params], dim=0)
        buf = ops.split3(w if w.stride(-1) == 1 else w.contiguous())
    if len(_pcache) > 1024:
        _pcache.clear()
    _pcache[key] = (ver, buf, tuple(weakref.ref(p) for p in params))
    return buf


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class Linear32Fn(Function):
    """y = act(x [W_1; ..; W_n]^T + [b_1; ..; b_n]) in fp32 mode (reference nn.Linear call sites; n > 1 = layers reading the
    same input, e.g. the two QueryPunish.query_weight projections, model/utils.py:100). x [..., K'] fp32 with K' >= K (zero
    padded), W_i fp32 [N_i, K]. act_grad_folded: the consumer's backward already returns d(pre-activation)."""

    @staticmethod
    def forward(ctx, x, act, folded, n, *wb):
        weights, biases = wb[:n], wb[n:]
        Kx = x.shape[-1]
        x2 = _c(x.reshape(-1, Kx))
        M, K = x2.shape[0], weights[0].shape[1]
        N = sum(w.shape[0] for w in weights)
        xp = ops.split3(x2)
        wp = weight_planes(list(weights), tag="lin")
        bias = None
        if biases[0] is not None:
            bias = biases[0].detach() if n == 1 else torch.cat([b.detach() for b in biases])
        y = torch.empty((M, N), dtype=F32, device=x.device)
        ops.gemm3(xp, 0, wp, 0, M, N, K, y, bias=bias, act=act)
        keep_y = act not in (None, "none") and not folded
        ctx.save_for_backward(xp, wp, y if keep_y else None)
        ctx.cfg = (act if keep_y else None, x.shape, [w.shape[0] for w in weights], K, biases[0] is not None, n)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        xp, wp, y = ctx.saved_tensors
        act, xshape, Ns, K, has_bias, n = ctx.cfg
        M, N, Kx = xp.shape[1], sum(Ns), xshape[-1]
        d = _c(dy.reshape(M, N))
        if act is not None:
            d = ops.act_bwd(d, y, act)
        dp = ops.split3(d)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, Kx), dtype=F32, device=d.device)
            ops.gemm3(dp, 0, wp, 1, M, Kx, N, dx)
            dx = dx.view(xshape)
        dw = torch.empty((N, K), dtype=F32, device=d.device)
        ops.gemm3(dp, 1, xp, 1, N, K, M, dw)
        db = ops.colsum(d) if has_bias else None
        gw, gb, r = [], [], 0
        for Ni in Ns:
            gw.append(dw[r:r + Ni])
            gb.append(db[r:r + Ni] if has_bias else None)
            r += Ni
        return (dx, None, None, None) + tuple(gw) + tuple(gb)


def linear32(x, weight, bias=None, act=None, act_grad_folded=False):
    return Linear32Fn.apply(x, act, act_grad_folded, 1, weight, bias)


def linear_cat32(x, w1, b1, w2, b2):
    return Linear32Fn.apply(x, None, False, 2, w1, w2, b1, b2)


# ---------------------------------------------------------------------------------------------------------------------
def _lstm32_args(S, H, T, D, gates, c_hist, seq_len):
    a = _lib.Lstm32Args()
    a.S, a.H, a.T, a.ndir = S, H, T, D
    a.gates, a.c_hist = gates.data_ptr(), c_hist.data_ptr()
    if seq_len is not None:
        a.seq_len = seq_len.data_ptr()
    return a


class Lstm32Fn(Function):
    """D directions of (bi)LSTMs over one input sequence, fp32 (reference nn.LSTM at model/Preprocessing.py:97-101,202).
    x [T, S, K'] fp32 time-major (K' >= K, zero padded); seq_len [S] int32 or None; per direction (w_ih, w_hh, b_ih, b_hh) in
    nn.LSTM layout; odd directions run backwards. Returns (seq_out [S, T, D*H] (zeros at padded positions) | None,
    h_last [S, D*H]: the state after the last valid position of each direction)."""

    @staticmethod
    def forward(ctx, x, seq_len, want_seq, *params):
        T, S, Kx = x.shape
        D = len(params) // 4
        w_ih, w_hh = [params[4 * d] for d in range(D)], [params[4 * d + 1] for d in range(D)]
        H, K = w_hh[0].shape[1], w_ih[0].shape[1]
        dev = x.device
        st = ops._stream()
        xp = ops.split3(_c(x).view(T * S, Kx))
        wih_p = weight_planes(w_ih, tag="lstm_ih")
        whh_p = torch.empty((D, 3, 4 * H, H), dtype=BF16, device=dev)
        with torch.no_grad():
            for d in range(D):
                ops.split3(w_hh[d].detach(), out=whh_p[d])
            bias = torch.cat([params[4 * d + 2].detach() + params[4 * d + 3].detach() for d in range(D)])
        gates = torch.empty((T * S, D * 4 * H), dtype=F32, device=dev)
        ops.gemm3(xp, 0, wih_p, 0, T * S, D * 4 * H, K, gates, bias=bias)
        h = torch.empty((D, S, H), dtype=F32, device=dev)
        c_hist = torch.empty((T + 1, D, S, H), dtype=F32, device=dev)
        h_planes = torch.empty((D, 3, S, H), dtype=BF16, device=dev)
        hprev_t = torch.empty((D, T, S, H), dtype=F32, device=dev)
        rec = torch.empty((D, S, 4 * H), dtype=F32, device=dev)
        seq_out = torch.empty((S, T, D * H), dtype=F32, device=dev) if want_seq else None
        h_last = torch.empty((S, D * H), dtype=F32, device=dev)
        a = _lstm32_args(S, H, T, D, gates, c_hist, seq_len)
        a.rec, a.h, a.h_planes, a.hprev_t = rec.data_ptr(), h.data_ptr(), h_planes.data_ptr(), hprev_t.data_ptr()
        a.h_last, a.h_last_ld = h_last.data_ptr(), D * H
        if seq_out is not None:
            a.seq_out, a.seq_out_ld = seq_out.data_ptr(), D * H
        for s in range(T):
            if s > 0:
                ops.gemm3(h_planes.view(D * 3, S, H), 0, whh_p.view(D * 3, 4 * H, H), 0, S, 4 * H, H, rec, batch=D,
                          c_batch=S * 4 * H, ldc=4 * H)
            a.s = s
            _lib.check(_lib.lstm32_cell_fwd(ctypes.byref(a), st), "dvgr_lstm32_cell_fwd")
        ctx.save_for_backward(xp, wih_p, whh_p, gates, c_hist, hprev_t, seq_len)
        ctx.cfg = (T, S, Kx, K, D, H)
        if want_seq:
            return seq_out, h_last
        return None, h_last

    @staticmethod
    def backward(ctx, d_seq, d_last):
        xp, wih_p, whh_p, gates, c_hist, hprev_t, seq_len = ctx.saved_tensors
        T, S, Kx, K, D, H = ctx.cfg
        dev = gates.device
        st = ops._stream()
        dh = torch.empty((D, S, H), dtype=F32, device=dev)
        dc = torch.empty((D, S, H), dtype=F32, device=dev)
        dgp = torch.empty((D, 3, S, 4 * H), dtype=BF16, device=dev)
        a = _lstm32_args(S, H, T, D, gates, c_hist, seq_len)
        dgates = torch.empty_like(gates)        # (the saved activated gates stay intact: the graph can be back-propagated again)
        a.dh, a.dc, a.dgate_planes, a.dgates = dh.data_ptr(), dc.data_ptr(), dgp.data_ptr(), dgates.data_ptr()
        if d_last is not None:
            d_last = _c(d_last)
            a.dh_last, a.dh_last_ld = d_last.data_ptr(), d_last.stride(0)
        if d_seq is not None:
            d_seq = _c(d_seq)
            a.dh_seq, a.dh_seq_ld = d_seq.data_ptr(), d_seq.stride(1)
        for s in range(T - 1, -1, -1):
            a.s = s
            _lib.check(_lib.lstm32_cell_bwd(ctypes.byref(a), st), "dvgr_lstm32_cell_bwd")
            if s > 0:        # dh_{s-1} += dgates_s W_hh (W_hh read MN-major)
                ops.gemm3(dgp.view(D * 3, S, 4 * H), 0, whh_p.view(D * 3, 4 * H, H), 1, S, H, 4 * H, dh, batch=D,
                          c_batch=S * H, ldc=H, beta=True)
        dgp_all = ops.split3(dgates)                              # [3, T*S, D*4H]: gate gradients of every step
        dwih = torch.empty((D * 4 * H, K), dtype=F32, device=dev)
        ops.gemm3(dgp_all, 1, xp, 1, D * 4 * H, K, T * S, dwih)
        db = ops.colsum(dgates)
        dwhh = torch.empty((D, 4 * H, H), dtype=F32, device=dev)
        for d in range(D):
            hp = ops.split3(hprev_t[d].view(T * S, H))
            ops.gemm(dgp_all, 1, hp, 1, 4 * H, H, 3 * ((T * S + 63) // 64) * 64, dwhh[d], k_inner=(T * S + 63) // 64,
                     a_c0=[d * 4 * H], a_c2=[0], a_c2_step=[1], b_c2=[2], b_c2_step=[-1])
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((T * S, Kx), dtype=F32, device=dev)
            ops.gemm3(dgp_all, 0, wih_p, 1, T * S, Kx, D * 4 * H, dx)
            dx = dx.view(T, S, Kx)
        grads = []
        for d in range(D):
            sl = slice(4 * H * d, 4 * H * (d + 1))
            grads += [dwih[sl], dwhh[d], db[sl], db[sl]]
        return (dx, None, None) + tuple(grads)


class Embed32Fn(Function):
    """tanh(dropout(embedding)) (reference model/Preprocessing.py:108-110) -> (words [B, L, Wp], time-major copy [L, B, Wp]),
    zero padded to Wp columns."""

    @staticmethod
    def forward(ctx, tokens, table, p):
        W = table.shape[1]
        Wp = (W + 7) // 8 * 8
        seed, sid = ag._site()
        words, x_tm = ops.embed_fwd(_c(tokens), table.detach(), Wp, p, seed, sid, out_dtype=F32)
        ctx.save_for_backward(tokens, words)
        ctx.cfg = (W, p, seed, sid, table.shape[0])
        return words, x_tm

    @staticmethod
    def backward(ctx, d_words, d_x):
        tokens, words = ctx.saved_tensors
        W, p, seed, sid, V = ctx.cfg
        dtable = torch.zeros((V, W), dtype=F32, device=words.device)
        ops.embed_bwd(_c(tokens), words, _c(d_words) if d_words is not None else None, _c(d_x) if d_x is not None else None,
                      W, dtable, p, seed, sid)
        return None, dtable, None


# ---------------------------------------------------------------------------------------------------------------------
class Gat32Fn(Function):
    """One punishGAT (reference model/GraphNN.py:95-113,174-178) in fp32: input dropout, the heads' projections as ONE split
    product, the fused attention kernel (fp32 build), output dropout inside it.
    x [B, N, D] fp32, gate [B, N] fp32, adj [N, N]; params per head: W.weight, W.bias, a.weight, a.bias."""

    @staticmethod
    def forward(ctx, p, heads, x, gate, adj, *params):
        B, N, D = x.shape
        M, Dh = B * N, D // heads
        seed, sid = ag._site(3)
        x2 = _c(x).view(M, D)
        xt = ops.dropout_raw(x2, p, seed, sid) if p > 0 else x2
        Ws = [params[4 * k] for k in range(heads)]
        wp = weight_planes(Ws, tag="gat")
        with torch.no_grad():
            bias = torch.cat([params[4 * k + 1].detach() for k in range(heads)])
            avec = torch.cat([torch.cat([params[4 * k + 2].detach().reshape(-1), params[4 * k + 3].detach().reshape(-1)])
                              for k in range(heads)]).view(heads, 2 * Dh + 1)
        xtp = ops.split3(xt)
        wh = torch.empty((M, D), dtype=F32, device=x.device)
        ops.gemm3(xtp, 0, wp, 0, M, D, D, wh, bias=bias)
        gate = _c(gate)
        outs, _ = ops.gat_attn_fwd([wh], [gate], [avec], adj, B, N, heads=heads, p_att=p, p_out=p, seed=seed, streams=[sid + 1])
        ctx.save_for_backward(xtp, wp, wh, outs[0], gate, avec, adj)
        ctx.cfg = (B, N, D, M, heads, Dh, p, seed, sid)
        return outs[0].view(B, N, D)

    @staticmethod
    def backward(ctx, dout):
        xtp, wp, wh, out, gate, avec, adj = ctx.saved_tensors
        B, N, D, M, heads, Dh, p, seed, sid = ctx.cfg
        dwhs, dgates, davecs = ops.gat_attn_bwd([wh], [gate], [avec], [out], [_c(dout).view(M, D)], adj, B, N, heads=heads,
                                                p_att=p, p_out=p, seed=seed, streams=[sid + 1])
        dwh = dwhs[0]
        dp = ops.split3(dwh)
        dx = None
        if ctx.needs_input_grad[2]:
            dxt = torch.empty((M, D), dtype=F32, device=dwh.device)
            ops.gemm3(dp, 0, wp, 1, M, D, D, dxt)
            dx = (ops.dropout_raw(dxt, p, seed, sid) if p > 0 else dxt).view(B, N, D)
        dW = torch.empty((D, D), dtype=F32, device=dwh.device)
        ops.gemm3(dp, 1, xtp, 1, D, D, M, dW)
        db = ops.colsum(dwh)
        dav = davecs[0]
        grads = []
        for k in range(heads):
            grads += [dW[k * Dh:(k + 1) * Dh], db[k * Dh:(k + 1) * Dh], dav[k, :2 * Dh].reshape(1, 2 * Dh), dav[k, 2 * Dh:].reshape(1)]
        return (None, None, dx, dgates[0], None) + tuple(grads)


def gat_layer32(gats, streams, xs, gates, adj, training):
    """fp32 counterpart of model.GraphNN.fused_gat_layer: one Gat32Fn per graph. Returns (per-stream stacks [n, B*N, D],
    per-graph outputs [B, N, D]) — in fp32 mode the dense outputs the auxiliary losses read ARE the graph outputs."""
    ns = max(streams) + 1
    outs = []
    for g, s in zip(gats, streams):
        p = g.dropout if training else 0.0
        outs.append(Gat32Fn.apply(float(p), g.n_heads, xs[s], gates[s], adj, *g.flat_params()))
    B, N, D = outs[0].shape
    stacks = [torch.stack([o.view(B * N, D) for o, s in zip(outs, streams) if s == k]) for k in range(ns)]
    return stacks, outs
