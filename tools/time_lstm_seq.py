"""Times the whole-sequence fused LSTM forward (dvgr_lstm_seq_fwd) against the two-kernel path (W_ih GEMM + T step
launches) at the workload shapes (appearance encoder: T=16, S=5120, K1=2048; question encoder: T=20, S=256, K1=304, 4 dirs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dualvgr_videoqa_b200.ops as ops

BF16 = torch.bfloat16


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for name, T, S, H, D, K1 in (("appearance", 16, 5120, 384, 2, 2048), ("question", 20, 256, 384, 4, 304)):
    x = (torch.randn(T, S, K1, device="cuda") * 0.5).to(BF16)
    wih = (torch.randn(D * 4 * H, K1, device="cuda") * 0.02).to(BF16)
    whh = (torch.randn(D, 4 * H, H, device="cuda") * 0.05).to(BF16)
    bias = torch.randn(D * 4 * H, device="cuda") * 0.1
    res = {}

    def fused():
        res["f"] = ops.lstm_seq_fwd(x, wih, whh, bias)

    def two():
        g = ops.linear_fwd(x.view(T * S, K1), wih, bias=bias, bn=256).view(T, S, D * 4 * H)
        res["t"] = ops.lstm_fwd(g, whh)

    mf, bf_ = timeit(fused)
    mt, bt = timeit(two)
    flops = 2.0 * S * D * 4 * H * (T * K1 + (T - 1) * H)
    err = int(res["f"][5][-1])
    d = float((res["f"][3].float() - res["t"][2].float()).norm() / res["t"][2].float().norm())
    print(f"{name}: fused {mf:.3f} ms (best {bf_:.3f}) = {flops / mf / 1e9:.0f} TFLOP/s | two-kernel {mt:.3f} ms (best {bt:.3f}) | "
          f"h_last rel diff {d:.2e} | timeouts {err}", flush=True)

    # backward: per-step launches vs the whole-sequence persistent launch (same operands; gates restored each run)
    gates_act, h_hist, c_hist = res["f"][0], res["f"][1], res["f"][2]
    dh = (torch.randn(S, D * H, device="cuda") * 0.1).to(BF16)
    g_std, c_std = ops.lstm_unblock_gates(gates_act, S).contiguous(), ops.lstm_unblock_c(c_hist, S).contiguous()
    work = g_std.clone()

    def bwd_steps():
        work.copy_(g_std)
        ops.lstm_bwd(work, whh, h_hist, c_std, dh)

    def bwd_seq():
        work.copy_(g_std)
        res["b"] = ops.lstm_bwd(gates_act, whh, h_hist, c_hist, dh, whole_sequence=True)

    def copy_only():
        work.copy_(g_std)

    mc, _ = timeit(copy_only)
    ms_, _ = timeit(bwd_steps)
    mq, _ = timeit(bwd_seq)
    print(f"{name}: backward per-step {ms_ - mc:.3f} ms | whole-sequence {mq - mc:.3f} ms | timeouts {int(res['b'][1][-1])}", flush=True)
