"""fp32 mode (north_star: "within 1e-4 relative error in fp32"): the 3 x bf16 split products on the tcgen05 GEMM, the fp32
variants of the fused kernels, the fp32 LSTM path, and the whole model against the float64 oracle at 1e-4."""
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


@pytest.mark.parametrize("M,N,K", [(300, 768, 768), (1000, 136, 300), (257, 3072, 2048)])
def test_split_products_reach_fp32_accuracy(ops, M, N, K):
    """x W^T, dy W and dy^T x from bf16 planes [lo | hi | hi]: a_lo b_hi + a_hi b_hi + a_hi b_lo in ONE launch each,
    against float64; a single bf16 pass sits at ~3e-3."""
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn((M, K), generator=g).cuda()
    w = (torch.randn((N, K), generator=g) * 0.05).cuda()
    b = torch.randn(N, generator=g).cuda()
    dy = torch.randn((M, N), generator=g).cuda()
    x3, w3, dy3 = ops.split3(x), ops.split3(w), ops.split3(dy)
    assert rel(x3[0].float()[:, :K] + x3[1].float()[:, :K], x) < 2e-5 and torch.equal(x3[1], x3[2])
    y = torch.empty((M, N), dtype=F32, device="cuda")
    ops.gemm3(x3, 0, w3, 0, M, N, K, y, bias=b, act="elu")
    ref = torch.nn.functional.elu(x.double() @ w.double().t() + b.double())
    assert rel(y, ref) < 2e-5, rel(y, ref)
    y1 = ops.linear_fwd(x.to(BF16) if K % 8 == 0 else torch.nn.functional.pad(x, (0, (-K) % 8)).to(BF16),
                        w.to(BF16) if K % 8 == 0 else torch.nn.functional.pad(w, (0, (-K) % 8)).to(BF16), bias=b, act="elu",
                        out_dtype=F32)
    assert rel(y1, ref) > 20 * rel(y, ref)                                    # the split really buys two orders of magnitude
    dx = torch.empty((M, K), dtype=F32, device="cuda")
    ops.gemm3(dy3, 0, w3, 1, M, K, N, dx)                                     # dgrad: W read MN-major
    assert rel(dx, dy.double() @ w.double()) < 2e-5
    dw = torch.zeros((N, K), dtype=F32, device="cuda")
    ops.gemm3(dy3, 1, x3, 1, N, K, M, dw)                                     # wgrad: both MN-major
    assert rel(dw, dy.double().t() @ x.double()) < 2e-5
    ops.gemm3(dy3, 1, x3, 1, N, K, M, dw, beta=True)                          # accumulate
    assert rel(dw, 2 * (dy.double().t() @ x.double())) < 2e-5
