"""CPU ORACLE for the DualVGR reasoning core — TEST INFRASTRUCTURE ONLY.

A functional, dtype-generic (run it in float64 for "truth") restatement of the reference's algorithm for the hot path
named in BASELINE.json. It takes a reference-compatible ``state_dict`` (plain ``{name: tensor}``) and reproduces
``DualVGR.forward`` plus the training loss of ``train.py``. Gradients come from torch autograd over these formulas.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this file; the product package (``dualvgr-videoqa_b200/``) never does.

PINNING: the reference ships no tests or golden vectors. This oracle is pinned against outputs of the
reference's own modules executed in the build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks it against those fixtures on every run.

Every function cites the reference lines it restates (paths relative to the reference root).
Dropout: every function takes optional multiplicative masks (already scaled by 1/(1-p)); ``None`` = dropout off /
eval mode, which is the configuration all parity fixtures use (fused kernels own their Philox stream).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- deterministic fixtures
def make_vocab(V, A):
    """Minimal vocab with the two tables DualVGR.__init__ reads (model/models.py:40-41)."""
    return {"question_token_to_idx": {f"w{i}": i for i in range(V)},
            "answer_token_to_idx": {f"a{i}": i for i in range(A)}}


def state_dict_spec(U, A, V, D=768, W=300, Dv=2048):
    """(name, shape) of every floating-point entry of the reference state_dict, in the reference's order
    (model/models.py:36-53 and the sub-module constructors; probed list in SURVEY.md §8b)."""
    H, K, Dh = D // 2, 4, D // 4
    spec = [("feature_aggregation.v_proj.weight", (D, D)), ("feature_aggregation.attn.weight", (1, D)),
            ("feature_aggregation.attn.bias", (1,)), ("linguistic_input_unit.encoder_embed.weight", (V, W))]
    for rnn in ("linguistic_input_unit.concatRNN.rnn", "linguistic_input_unit.encoder"):
        for sfx in ("", "_reverse"):
            spec += [(f"{rnn}.weight_ih_l0{sfx}", (4 * H, W)), (f"{rnn}.weight_hh_l0{sfx}", (4 * H, H)),
                     (f"{rnn}.bias_ih_l0{sfx}", (4 * H,)), (f"{rnn}.bias_hh_l0{sfx}", (4 * H,))]
    for sfx in ("", "_reverse"):
        e = "visual_appearance_input_unit.encoder"
        spec += [(f"{e}.weight_ih_l0{sfx}", (4 * H, Dv)), (f"{e}.weight_hh_l0{sfx}", (4 * H, H)),
                 (f"{e}.bias_ih_l0{sfx}", (4 * H,)), (f"{e}.bias_hh_l0{sfx}", (4 * H,))]
    spec += [("visual_motion_input_unit.weight", (D, Dv)), ("visual_motion_input_unit.bias", (D,))]
    u = "visual_input_unit"
    for i in range(U):
        spec += [(f"{u}.queryAttn.{i}.feat_enhance.weight", (D, D)), (f"{u}.queryAttn.{i}.feat_enhance.bias", (D,)),
                 (f"{u}.queryAttn.{i}.fc.weight", (1, D)), (f"{u}.queryAttn.{i}.fc.bias", (1,))]
    for name in ("queryPunish_appear", "queryPunish_motion"):
        for i in range(U):
            spec += [(f"{u}.{name}.{i}.query_weight.weight", (D, W)), (f"{u}.{name}.{i}.query_weight.bias", (D,))]
    for name in ("appearance_GCN", "motion_GCN", "acGCN", "mcGCN"):
        for i in range(U):
            for k in range(K):
                p = f"{u}.{name}.{i}.attention_{k}"
                spec += [(f"{p}.W.weight", (Dh, D)), (f"{p}.W.bias", (Dh,)), (f"{p}.a.weight", (1, 2 * Dh)),
                         (f"{p}.a.bias", (1,))]
    for name in ("attention_appearance", "attention_motion"):
        for i in range(U):
            p = f"{u}.{name}.{i}.project"
            spec += [(f"{p}.0.weight", (D, D)), (f"{p}.0.bias", (D,)), (f"{p}.2.weight", (1, D))]
    f = f"{u}.visualfusion"
    spec += [(f"{f}.linear0.weight", (512, D)), (f"{f}.linear0.bias", (512,)), (f"{f}.linear1.weight", (512, D)),
             (f"{f}.linear1.bias", (512,)), (f"{f}.linear_out.weight", (D, 256)), (f"{f}.linear_out.bias", (D,))]
    o = "output_unit"
    spec += [(f"{o}.question_proj.weight", (D, D)), (f"{o}.question_proj.bias", (D,)),
             (f"{o}.classifier.1.weight", (D, 2 * D)), (f"{o}.classifier.1.bias", (D,)),
             (f"{o}.classifier.3.weight", (D,)), (f"{o}.classifier.3.bias", (D,)),
             (f"{o}.classifier.3.running_mean", (D,)), (f"{o}.classifier.3.running_var", (D,)),
             (f"{o}.classifier.5.weight", (A, D)), (f"{o}.classifier.5.bias", (A,))]
    return spec


def make_state_dict(U, A, V, seed=666, D=768, W=300, Dv=2048):
    """Deterministic, construction-order-independent weights: each tensor is drawn from its own CPU generator seeded by
    (seed, index) — xavier-uniform-sized matrices (the reference re-initialises every Linear/LSTM that way,
    model/utils.py:8-33), N(0, 0.02) biases (zero biases make QueryAttn's normalize singular), BN affine
    near (1, 0), non-trivial running stats. Regenerated identically on any machine: no weight file is ever shipped."""
    sd = {}
    for idx, (name, shape) in enumerate(state_dict_spec(U, A, V, D, W, Dv)):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if name.endswith("running_var"):
            t = 0.5 + torch.rand(shape, generator=g, dtype=torch.float32)
        elif name.endswith("running_mean"):
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif name.endswith("classifier.3.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float32)
        elif name.endswith("encoder_embed.weight"):
            t = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1                      # model/models.py:53
            t[0].zero_()
        elif len(shape) == 2 and shape[0] > 1:
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        elif len(shape) == 2:                                               # [1, D] scoring vectors
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound
        else:
            t = 0.02 * torch.randn(shape, generator=g, dtype=torch.float32)
        sd[name] = t.float()
    sd["output_unit.classifier.3.num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)
    return sd


def make_inputs(B, N, L, A, V, seed=1, F_=16, Dv=2048):
    """Synthetic batch with the shapes/dtypes of the reference loader (DataLoader.py:61-84):
    |N(0,1)| features, question_len ~ U{3..L} with row 0 full-length, tokens U{2..V-1}, zero padding."""
    g = torch.Generator().manual_seed(seed)
    app = torch.randn((B, N, F_, Dv), generator=g, dtype=torch.float32).abs_()
    mot = torch.randn((B, N, Dv), generator=g, dtype=torch.float32).abs_()
    qlen = torch.randint(min(3, L), L + 1, (B,), generator=g)
    qlen[0] = L
    q = torch.randint(2, V, (B, L), generator=g)
    q = q * (torch.arange(L)[None, :] < qlen[:, None])
    ans = torch.randint(0, A, (B,), generator=g)
    return app, mot, q.long(), qlen.long(), ans.long()


# ----------------------------------------------------------------------------------------------- building blocks
def build_adjacency(N):
    """model/models.py:114-119 (+ normalize :26-33): ones(N,N) symmetrised, + I, row-normalised.
    Result: 2/(N+1) on the diagonal, 1/(N+1) elsewhere."""
    a = np.ones((N, N), dtype=np.float64) + np.eye(N)
    a = a / a.sum(1, keepdims=True)
    return torch.from_numpy(a.astype(np.float32))


def linear(x, w, b=None):
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, reverse=False, lengths=None):
    """One direction of nn.LSTM (gate order i, f, g, o) written out step by step.
    x [S, T, In]; lengths [S] or None. With lengths, padded steps carry the state and output zeros, which is what
    pack_padded_sequence / pad_packed_sequence produce (model/Preprocessing.py:26-36,119-121).
    Returns (outputs [S, T, H], final hidden [S, H])."""
    S, T, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros((S, H))
    c = x.new_zeros((S, H))
    outs = [None] * T
    gx = linear(x, w_ih, b_ih + b_hh)
    for s in range(T):
        t = T - 1 - s if reverse else s
        pre = gx[:, t] + h @ w_hh.t()
        i, f, g, o = pre.split(H, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        if lengths is not None:
            live = (t < lengths).to(x.dtype).unsqueeze(1)
            c = live * c_new + (1 - live) * c
            h = live * h_new + (1 - live) * h
            outs[t] = live * h_new
        else:
            c, h = c_new, h_new
            outs[t] = h_new
    return torch.stack(outs, dim=1), h


def bilstm(sd, prefix, x, lengths=None):
    """Bidirectional single-layer LSTM with the reference's parameter names."""
    of, hf = lstm_direction(x, sd[f"{prefix}.weight_ih_l0"], sd[f"{prefix}.weight_hh_l0"], sd[f"{prefix}.bias_ih_l0"],
                            sd[f"{prefix}.bias_hh_l0"], False, lengths)
    ob, hb = lstm_direction(x, sd[f"{prefix}.weight_ih_l0_reverse"], sd[f"{prefix}.weight_hh_l0_reverse"],
                            sd[f"{prefix}.bias_ih_l0_reverse"], sd[f"{prefix}.bias_hh_l0_reverse"], True, lengths)
    return torch.cat([of, ob], dim=-1), torch.cat([hf, hb], dim=-1)


def question_encoder(sd, question, question_len, emb_mask=None, final_mask=None):
    """InputUnitLinguisticDynamic.forward, model/Preprocessing.py:106-127 (DynamicRNN :17-45).
    Returns (question_embedding [B,D], words [B,L,W], dynamic_q [B,L,D])."""
    p = "linguistic_input_unit"
    e = sd[f"{p}.encoder_embed.weight"][question]
    if emb_mask is not None:
        e = e * emb_mask
    words = torch.tanh(e)
    dynamic_q, _ = bilstm(sd, f"{p}.concatRNN.rnn", words, question_len)       # zero rows at padded steps (:35-36)
    _, q_emb = bilstm(sd, f"{p}.encoder", words, question_len)                 # final states of the packed run
    if final_mask is not None:
        q_emb = q_emb * final_mask
    return q_emb, words, dynamic_q


def appearance_encoder(sd, app, in_mask=None, out_mask=None):
    """VisualAppearanceEncoder.forward, model/Preprocessing.py:209-234: tanh(dropout(x)) -> BiLSTM over the F frames of
    every clip -> cat(final fwd, final bwd) -> dropout. app [B,N,F,Dv] -> [B,N,D]."""
    B, N, F_, Dv = app.shape
    x = app if in_mask is None else app * in_mask
    x = torch.tanh(x).reshape(B * N, F_, Dv)
    _, h = bilstm(sd, "visual_appearance_input_unit.encoder", x)
    if out_mask is not None:
        h = h * out_mask.reshape(B * N, -1)
    return h.reshape(B, N, -1)


def query_attn(sd, i, words, dynamic_q, question_len):
    """QueryAttn.forward, model/utils.py:66-84. The softmax runs over ALL L positions, padding included; the mask and
    the renormalisation (eps 1e-5) come after. Returns (q_c [B,W], alpha [B,L])."""
    p = f"visual_input_unit.queryAttn.{i}"
    y = linear(dynamic_q, sd[f"{p}.feat_enhance.weight"], sd[f"{p}.feat_enhance.bias"])
    d = y / y.norm(dim=-1, keepdim=True).clamp_min(1e-12)                      # F.normalize(p=2, eps=1e-12)
    score = linear(d, sd[f"{p}.fc.weight"], sd[f"{p}.fc.bias"]).squeeze(-1)
    alpha = torch.softmax(score, dim=1)
    L = alpha.shape[1]
    mask = (torch.arange(L, device=alpha.device)[None, :] < question_len[:, None]).to(alpha.dtype)   # :72-75
    alpha = alpha * mask
    alpha = alpha / (alpha.sum(1, keepdim=True) + 1e-5)
    q_c = torch.einsum("bl,blw->bw", alpha, words)
    return q_c, alpha


def query_punish(sd, name, i, q_c, X):
    """QueryPunish.forward, model/utils.py:92-105: one sigmoid gate per clip (the reference returns it expanded to Dh
    columns; the scalar per node is the information content). Returns g [B,N]."""
    p = f"visual_input_unit.{name}.{i}.query_weight"
    query = linear(q_c, sd[f"{p}.weight"], sd[f"{p}.bias"])
    return torch.sigmoid(torch.einsum("bnd,bd->bn", X, query))


def punish_gat(sd, name, i, x, adj, gate, in_mask=None, att_masks=None, out_mask=None, heads=4, slope=0.01):
    """punishGAT.forward (model/GraphNN.py:174-178) over PunishGraphAttentionLayer.forward (:95-113).
    The pairwise logit a.[Wh_i ; Wh_j] + c of :115-155 is evaluated as a[:Dh].Wh_i + a[Dh:].Wh_j + c (same value
    without the [B,N,N,2Dh] tensor). The gate multiplies the VALUES only, after the logits (:98-104)."""
    if in_mask is not None:
        x = x * in_mask
    outs = []
    for k in range(heads):
        p = f"visual_input_unit.{name}.{i}.attention_{k}"
        Wh = linear(x, sd[f"{p}.W.weight"], sd[f"{p}.W.bias"])                 # [B,N,Dh]
        a = sd[f"{p}.a.weight"][0]
        Dh = Wh.shape[-1]
        s = Wh @ a[:Dh]
        t = Wh @ a[Dh:]
        e = F.leaky_relu(s[:, :, None] + t[:, None, :] + sd[f"{p}.a.bias"], slope)
        e = torch.where(adj > 0, e, torch.full_like(e, -9e15))
        P = torch.softmax(e, dim=-1)
        if att_masks is not None:
            P = P * att_masks[k]
        V = Wh * gate[:, :, None]
        outs.append(F.elu(P @ V))
    out = torch.cat(outs, dim=2)
    if out_mask is not None:
        out = out * out_mask
    return out


def attention_sfgcn(sd, name, i, z_common, z_specific):
    """AttentionSFGCN.forward, model/Attention.py:20-23, on the stack built at model/models.py:163-166.
    Returns (embed [B,N,D], beta [B,2,N])."""
    p = f"visual_input_unit.{name}.{i}.project"
    z = torch.stack([z_common, z_specific], dim=1)
    w = linear(torch.tanh(linear(z, sd[f"{p}.0.weight"], sd[f"{p}.0.bias"])), sd[f"{p}.2.weight"]).squeeze(-1)
    beta = torch.softmax(w, dim=1)
    return (beta.unsqueeze(-1) * z).sum(1), beta


def mfb(sd, a, m):
    """MFB.forward with the constructor arguments of model/models.py:109 (mm_dim 256, factor 2, ELU in/out, no
    dropout, no normalisation): model/fusions/fusions.py:419-453."""
    p = "visual_input_unit.visualfusion"
    x0 = F.elu(linear(a, sd[f"{p}.linear0.weight"], sd[f"{p}.linear0.bias"]))
    x1 = F.elu(linear(m, sd[f"{p}.linear1.weight"], sd[f"{p}.linear1.bias"]))
    z = (x0 * x1).reshape(*x0.shape[:-1], 256, 2).sum(-1)
    return F.elu(linear(z, sd[f"{p}.linear_out.weight"], sd[f"{p}.linear_out.bias"]))


def context_self_attn(sd, v, mask=None):
    """ContextSelfAttn.forward, model/AnswerDecoder.py:165-182: the DROPPED features are both scored and pooled."""
    p = "feature_aggregation"
    if mask is not None:
        v = v * mask
    score = linear(F.elu(linear(v, sd[f"{p}.v_proj.weight"])), sd[f"{p}.attn.weight"], sd[f"{p}.attn.bias"])
    alpha = torch.softmax(score, dim=1)
    return (alpha * v).sum(1)


def output_unit(sd, q_emb, v, training, mask1=None, mask2=None, eps=1e-5):
    """SimpleOutputUnitOpenEnded.forward, model/AnswerDecoder.py:197-202 (layers :188-195). BatchNorm1d uses batch
    statistics (biased variance) in training mode and the running statistics in eval mode."""
    p = "output_unit"
    q = linear(q_emb, sd[f"{p}.question_proj.weight"], sd[f"{p}.question_proj.bias"])
    x = torch.cat([v, q], dim=1)
    if mask1 is not None:
        x = x * mask1
    x = F.elu(linear(x, sd[f"{p}.classifier.1.weight"], sd[f"{p}.classifier.1.bias"]))
    if training:
        mean, var = x.mean(0), x.var(0, unbiased=False)
    else:
        mean, var = sd[f"{p}.classifier.3.running_mean"], sd[f"{p}.classifier.3.running_var"]
    x = (x - mean) / torch.sqrt(var + eps) * sd[f"{p}.classifier.3.weight"] + sd[f"{p}.classifier.3.bias"]
    if mask2 is not None:
        x = x * mask2
    return linear(x, sd[f"{p}.classifier.5.weight"], sd[f"{p}.classifier.5.bias"])


def dualvgr_unit_stack(sd, U, app, mot, dynamic_q, words, question_len, adj):
    """DualVGRUnit_multiple.forward, model/models.py:121-173 with graph_layers = 1 (all shipped configs)."""
    com_app_l, com_mot_l, aq_l, mq_l = [], [], [], []
    aq_embed = mq_embed = None
    for i in range(U):
        q_c, _ = query_attn(sd, i, words, dynamic_q, question_len)
        g_a = query_punish(sd, "queryPunish_appear", i, q_c, app)
        g_m = query_punish(sd, "queryPunish_motion", i, q_c, mot)
        com_app = punish_gat(sd, "acGCN", i, app, adj, g_a)
        aq = punish_gat(sd, "appearance_GCN", i, app, adj, g_a)
        com_mot = punish_gat(sd, "mcGCN", i, mot, adj, g_m)
        mq = punish_gat(sd, "motion_GCN", i, mot, adj, g_m)
        com_app_l.append(com_app); com_mot_l.append(com_mot); aq_l.append(aq); mq_l.append(mq)
        aq_embed, _ = attention_sfgcn(sd, "attention_appearance", i, com_app, aq)
        mq_embed, _ = attention_sfgcn(sd, "attention_motion", i, com_mot, mq)
        app = app + aq_embed
        mot = mot + mq_embed
    visual = mfb(sd, app, mot)
    return visual, aq_embed, mq_embed, com_app_l, com_mot_l, aq_l, mq_l


def dualvgr_forward(sd, U, app, mot, question, question_len, training=True, adj=None):
    """DualVGR.forward, model/models.py:55-83, dropout off. Returns the reference's 7-tuple."""
    N = app.shape[1]
    if adj is None:
        adj = build_adjacency(N).to(app.dtype)
    q_emb, words, dynamic_q = question_encoder(sd, question, question_len)
    a = appearance_encoder(sd, app)
    m = linear(mot, sd["visual_motion_input_unit.weight"], sd["visual_motion_input_unit.bias"])
    visual, aq_embed, mq_embed, ca, cm, aq, mq = dualvgr_unit_stack(sd, U, a, m, dynamic_q, words, question_len, adj)
    pooled = context_self_attn(sd, visual)
    logits = output_unit(sd, q_emb, pooled, training)
    return logits, aq_embed, mq_embed, ca, cm, aq, mq


# ----------------------------------------------------------------------------------------------- losses
def common_loss(e1, e2):
    """utils.py:10-18: centre over nodes, L2-normalise over features, squared difference of the two N x N Grams,
    MEAN over [B,N,N]."""
    def prep(e):
        e = e - e.mean(dim=1, keepdim=True)
        return e / e.norm(dim=2, keepdim=True).clamp_min(1e-12)
    a, b = prep(e1), prep(e2)
    return ((a @ a.transpose(1, 2) - b @ b.transpose(1, 2)) ** 2).mean()


def loss_dependence(e1, e2, dim):
    """utils.py:20-31 (HSIC): sum over the batch of trace(R K1 R K2), R = I - 1/dim."""
    R = torch.eye(dim, dtype=e1.dtype, device=e1.device) - 1.0 / dim
    K1 = e1 @ e1.transpose(1, 2)
    K2 = e2 @ e2.transpose(1, 2)
    return torch.einsum("bij,bji->", R @ K1, R @ K2)


def train_loss(outputs, answers, N, alpha=1.0, beta=1e-8):
    """train.py:146-154: CE + alpha * mean_l common_loss + beta * mean_l (HSIC_app + HSIC_mot).
    Returns (total, ce, loss_com_sum, loss_dep_sum)."""
    logits, _, _, ca, cm, aq, mq = outputs
    ce = F.cross_entropy(logits, answers)
    dep = sum(loss_dependence(aq[i], ca[i], N) + loss_dependence(mq[i], cm[i], N) for i in range(len(aq)))
    com = sum(common_loss(ca[i], cm[i]) for i in range(len(aq)))
    n = len(aq)
    return ce + alpha * com / n + beta * dep / n, ce, com, dep


def cast_state_dict(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
