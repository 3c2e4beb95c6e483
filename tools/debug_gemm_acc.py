import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dualvgr_videoqa_b200.ops as ops
BF = torch.bfloat16
g = torch.Generator().manual_seed(0)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for M2 in (200, 256, 10240):
    D = 768
    dh = (torch.randn((2, M2, D), generator=g) * 0.1).to(BF).cuda()
    w = (torch.randn((2, D, D), generator=g) * 0.05).to(BF).cuda()
    c0 = (torch.randn((2, M2, D), generator=g) * 0.1).to(BF).cuda()
    c = c0.clone()
    ops.gemm(dh, 0, w, 1, M2, D, D, c, ldc=D, beta=True, batch=2, c_batch=M2 * D, a_c2=[0, 1], b_c2=[0, 1])
    ref = c0.float() + torch.einsum("bmn,bnk->bmk", dh.float(), w.float())
    c1 = c0.clone()
    ops.gemm(dh, 0, w, 1, M2, D, D, c1, ldc=D, beta=False, batch=2, c_batch=M2 * D, a_c2=[0, 1], b_c2=[0, 1])
    print(M2, "acc", rel(c, ref), rel(c[0], ref[0]), rel(c[1], ref[1]), "noacc", rel(c1, ref - c0.float()))
    # 4-batch variant with stacked A [2,M,D] read twice each
    x = (torch.randn((2, M2, D), generator=g) * 0.1).to(BF).cuda()
    w4 = (torch.randn((4, D, D), generator=g) * 0.05).to(BF).cuda()
    bias = torch.randn((4, D), generator=g).cuda()
    out = torch.empty((4, M2, D), dtype=BF, device="cuda")
    ops.gemm(x, 0, w4, 0, M2, D, D, out, ldc=D, bias=bias, batch=4, c_batch=M2 * D, bias_batch=D, a_c2=[0, 0, 1, 1], b_c2=[0, 1, 2, 3])
    ref4 = torch.einsum("gmk,gnk->gmn", x.float()[[0, 0, 1, 1]], w4.float()) + bias[:, None, :]
    print("   batch4", [round(rel(out[i], ref4[i]), 5) for i in range(4)])
