"""Mirror of the live part of reference model/Attention.py: AttentionSFGCN (:11-23), the two-view fusion."""
import torch
import torch.nn as nn

from dualvgr_videoqa_b200 import autograd as ag

BF16 = torch.bfloat16


class AttentionSFGCN(nn.Module):
    def __init__(self, in_size, hidden_size=16):
        super().__init__()
        self.project = nn.Sequential(nn.Linear(in_size, hidden_size), nn.Tanh(), nn.Linear(hidden_size, 1, bias=False))

    def fused(self, z_stack, x):
        """z_stack [2, M, D] bf16 (common, specific), x [B,N,D] bf16 -> (x + embed, embed): GEMM with tanh epilogue, then
        the fused softmax-over-views / weighted sum / residual kernel."""
        hidden = ag.linear(z_stack, self.project[0].weight, self.project[0].bias, act="tanh", act_grad_folded=True)
        return ag.ViewAttnFn.apply(hidden, z_stack, x, self.project[2].weight)

    def forward(self, z):
        """z [B, 2, N, D] -> ((beta * z).sum(1) [B,N,D], beta [B,2,N,1]) as the reference (:20-23)."""
        if z.dim() != 4 or z.shape[1] != 2:
            raise NotImplementedError("AttentionSFGCN on sm_100a fuses exactly two views (reference model/models.py:163-166)")
        B, _, N, D = z.shape
        zs = z.to(ag.ACT[0]).transpose(0, 1).reshape(2, B * N, D)
        zero = torch.zeros((B, N, D), dtype=ag.ACT[0], device=z.device)
        _, embed = self.fused(zs, zero)
        hidden = ag.linear(zs, self.project[0].weight, self.project[0].bias, act="tanh")
        w = (hidden.float() @ self.project[2].weight.float().t()).view(2, B, N, 1).transpose(0, 1)
        return embed.to(z.dtype), torch.softmax(w, dim=1).to(z.dtype)
