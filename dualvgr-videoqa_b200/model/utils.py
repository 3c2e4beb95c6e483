"""Mirror of the live part of reference model/utils.py: init_modules (:8-33), QueryAttn (:60-84), QueryPunish (:86-105).
The two modules keep the reference's parameter names; their forward runs the fused Query Punishment kernels."""
import numpy as np
import torch
import torch.nn as nn
from torch.nn import init

from dualvgr_videoqa_b200 import autograd as ag

BF16 = torch.bfloat16
_INITS = {"normal": init.normal_, "xavier_normal": init.xavier_normal_, "xavier_uniform": init.xavier_uniform_,
          "kaiming_normal": init.kaiming_normal_, "kaiming_uniform": init.kaiming_uniform_,
          "orthogonal": init.orthogonal_}


def init_modules(modules, w_init='kaiming_uniform'):
    """Re-initialises every Linear / Conv (zero bias) and every LSTM / GRU (zero biases) — reference model/utils.py:8-33."""
    if w_init not in _INITS:
        raise NotImplementedError
    fn = _INITS[w_init]
    for m in modules:
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
            fn(m.weight)
            if m.bias is not None:
                init.zeros_(m.bias)
        if isinstance(m, (nn.LSTM, nn.GRU)):
            for name, param in m.named_parameters():
                if 'bias' in name:
                    init.zeros_(param)
                elif 'weight' in name:
                    fn(param)


def pad_last(x, mult=8):
    """Zero-pads the last dim to a multiple of `mult` (TMA needs 16-byte row pitches: word_dim 300 -> 304)."""
    k = x.shape[-1]
    kp = (k + mult - 1) // mult * mult
    return x if kp == k else torch.nn.functional.pad(x, (0, kp - k))


class QueryAttn(nn.Module):
    def __init__(self, module_dim=768):
        super().__init__()
        self.feat_enhance = nn.Linear(module_dim, module_dim)
        self.fc = nn.Linear(module_dim, 1)

    def forward(self, word_embedding, dynamic_question_embedding, question_len, word_dim=None):
        """word_embedding [B,L,W], dynamic_question_embedding [B,L,D], question_len [B] -> (q_c [B,W], attn [B,L]).
        word_dim given (internal fast path): inputs are already bf16, words K-padded, question_len int32, and q_c is
        returned as the padded bf16 [B, Wp] operand of the QueryPunish projections."""
        fast = word_dim is not None
        W = word_dim if fast else word_embedding.shape[-1]
        words = word_embedding if fast else pad_last(word_embedding).to(ag.ACT[0])
        dq = dynamic_question_embedding if fast else dynamic_question_embedding.to(ag.ACT[0])
        qlen = question_len if fast else question_len.to(torch.int32)
        y = ag.linear(dq, self.feat_enhance.weight, self.feat_enhance.bias)
        qc, attn = ag.QAttnFn.apply(y, words, qlen, self.fc.weight, self.fc.bias, W)
        if fast:
            return qc, attn
        return qc[:, :W].to(word_embedding.dtype), attn


class QueryPunish(nn.Module):
    def __init__(self, word_dim=300, module_dim=768):
        super().__init__()
        self.temp = np.sqrt(word_dim * module_dim)
        self.query_weight = nn.Linear(word_dim, module_dim)

    def query(self, question_guided_padded):
        return ag.linear(question_guided_padded, self.query_weight.weight, self.query_weight.bias)

    def forward(self, question_guided, visual_feature):
        """question_guided [B,W], visual_feature [B,N,D] -> scores [B,N,D/4] (a stride-0 expansion of [B,N,1], as in the
        reference :103)."""
        q = self.query(pad_last(question_guided).to(ag.ACT[0]))
        x = visual_feature.to(ag.ACT[0])
        g, _ = ag.GateFn.apply(x, x, torch.cat([q, q], dim=1))
        g = g.to(visual_feature.dtype).unsqueeze(-1)
        return g.expand(g.size(0), g.size(1), visual_feature.size(2) // 4)
