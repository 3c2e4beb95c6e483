"""Mirror of the live part of reference model/AnswerDecoder.py: ContextSelfAttn (:155-182) and
SimpleOutputUnitOpenEnded (:184-202)."""
import torch
import torch.nn as nn

from dualvgr_videoqa_b200 import autograd as ag

BF16 = torch.bfloat16


class ContextSelfAttn(nn.Module):
    def __init__(self, module_dim=768):
        super().__init__()
        self.module_dim = module_dim
        self.v_proj = nn.Linear(module_dim, module_dim, bias=False)
        self.attn = nn.Linear(module_dim, 1)
        self.activation = nn.ELU()
        self.dropout = nn.Dropout(0.15)

    def forward(self, visual_feat):
        """[B,N,D] -> [B,D]: the dropped-out features are both scored and pooled (reference :173-180)."""
        dt = visual_feat.dtype
        v = ag.dropout(visual_feat.to(ag.ACT[0]), self.dropout.p, self.training)
        u = ag.linear(v, self.v_proj.weight, None, act="elu", act_grad_folded=True)
        pooled = ag.ReadoutFn.apply(v, u, self.attn.weight, self.attn.bias)
        return pooled if dt == pooled.dtype else pooled.to(dt)


class SimpleOutputUnitOpenEnded(nn.Module):
    def __init__(self, module_dim=512, num_answers=1000):
        super().__init__()
        self.question_proj = nn.Linear(module_dim, module_dim)
        self.classifier = nn.Sequential(nn.Dropout(0.15), nn.Linear(module_dim * 2, module_dim), nn.ELU(),
                                        nn.BatchNorm1d(module_dim), nn.Dropout(0.15), nn.Linear(module_dim, num_answers))

    def forward(self, question_embedding, visual_embedding):
        """([B,D], [B,D]) -> logits [B,A] fp32 (reference :197-202)."""
        c = self.classifier
        q = ag.linear(question_embedding.to(ag.ACT[0]), self.question_proj.weight, self.question_proj.bias)
        x = torch.cat([visual_embedding.to(ag.ACT[0]), q], dim=1)
        x = ag.dropout(x, c[0].p, self.training)
        # fp32 out: BatchNorm centres its input, which would amplify bf16 rounding of x by |x| / std
        x = ag.linear(x, c[1].weight, c[1].bias, act="elu", out_f32=True)
        bn = c[3]
        use_batch_stats = self.training or not bn.track_running_stats
        if self.training and bn.track_running_stats:
            bn.num_batches_tracked += 1
        grp = getattr(self, "sync_bn_group", None)          # set by engine.TrainEngine(sync_bn=True): global batch statistics
        sync = (grp, getattr(self, "sync_bn_world", 1)) if (grp is not None and self.training) else None
        x = ag.BatchNormFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, use_batch_stats,
                                 bn.momentum if bn.momentum is not None else 0.1, bn.eps, sync)
        x = ag.dropout(x, c[4].p, self.training)
        return ag.linear(x, c[5].weight, c[5].bias, out_f32=True)
