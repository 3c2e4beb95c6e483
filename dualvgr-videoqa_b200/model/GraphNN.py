"""Mirror of the live part of reference model/GraphNN.py: PunishGraphAttentionLayer (:77-155) and punishGAT (:159-178).
Parameter names are the reference's (attention_k.W / attention_k.a); the forward of a whole multi-head GAT is one
batched tcgen05 projection GEMM + one fused attention launch (csrc/gat.cu)."""
import torch
import torch.nn as nn

from dualvgr_videoqa_b200 import autograd as ag

BF16 = torch.bfloat16


class PunishGraphAttentionLayer(nn.Module):
    """One attention head: parameter container (W: in->out, a: 2*out->1). The arithmetic is fused across the heads of
    the owning punishGAT; a head is never evaluated on its own on the fast path."""

    def __init__(self, in_features, out_features, dropout, alpha, concat=True):
        super().__init__()
        if not concat:
            raise NotImplementedError("DualVGR only builds concat=True heads (reference model/GraphNN.py:168)")
        self.dropout, self.in_features, self.out_features, self.alpha, self.concat = dropout, in_features, out_features, alpha, concat
        self.W = nn.Linear(in_features, out_features)
        nn.init.xavier_uniform_(self.W.weight, gain=1.414)
        self.a = nn.Linear(2 * out_features, 1)
        nn.init.xavier_uniform_(self.a.weight, gain=1.414)
        self.leakyrelu = nn.LeakyReLU(self.alpha)

    def params(self):
        return [self.W.weight, self.W.bias, self.a.weight, self.a.bias]


def gate_from_scores(scores, B, N, device):
    """The reference passes QueryPunish's gate expanded to [B,N,Dh] (model/utils.py:103); the kernels take [B,N]."""
    if scores is None:
        return torch.ones((B, N), dtype=torch.float32, device=device)
    return (scores[..., 0] if scores.dim() == 3 else scores).float().contiguous()


class punishGAT(nn.Module):
    def __init__(self, n_feat, n_hid, dropout, alpha, n_heads, q_attn=True):
        super().__init__()
        self.dropout, self.q_attn, self.n_heads, self.alpha = dropout, q_attn, n_heads, alpha
        self.attentions = [PunishGraphAttentionLayer(n_feat, n_hid, dropout=dropout, alpha=alpha, concat=True)
                           for _ in range(n_heads)]
        for i, attention in enumerate(self.attentions):
            self.add_module('attention_{}'.format(i), attention)

    def flat_params(self):
        return [p for att in self.attentions for p in att.params()]

    def forward(self, x, adj, scores):
        """x [B,N,F], adj [N,N], scores [B,N,Dh] (or [B,N]) -> [B,N,heads*Dh]; same dtype as x."""
        B, N, _ = x.shape
        gate = gate_from_scores(scores, B, N, x.device)
        outs = ag.GatLayerFn.apply((0,), adj.float().contiguous(), self.dropout, self.training, self.n_heads,
                                   x.to(ag.ACT[0]), gate, *self.flat_params())
        return outs[0][0].view(B, N, -1).to(x.dtype)


def fused_gat_layer(gats, streams, xs, gates, adj, training):
    """Runs several punishGATs (<= 4) with one attention launch. gats[i] reads xs[streams[i]] / gates[streams[i]].
    Returns (per-stream stacks [n, B*N, D] bf16, per-graph fp32 [B,N,D])."""
    if xs[0].dtype == torch.float32:
        from dualvgr_videoqa_b200 import fp32_path
        return fp32_path.gat_layer32(gats, streams, xs, gates, adj, training)
    ns = max(streams) + 1
    params = [p for g in gats for p in g.flat_params()]
    outs = ag.GatLayerFn.apply(tuple(streams), adj, gats[0].dropout, training, gats[0].n_heads, *xs, *gates, *params)
    return outs[:ns], outs[ns:]
