"""Launches each whole-sequence LSTM kernel twice at the workload shapes (for `ncu -k regex:lstm_seq`):
appearance fwd, appearance bwd, question fwd, question bwd."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dualvgr_videoqa_b200.ops as ops

BF16 = torch.bfloat16
for name, T, S, H, D, K1 in (("appearance", 16, 5120, 384, 2, 2048), ("question", 20, 256, 384, 4, 304)):
    x = (torch.randn(T, S, K1, device="cuda") * 0.5).to(BF16)
    wih = (torch.randn(D * 4 * H, K1, device="cuda") * 0.02).to(BF16)
    whh = (torch.randn(D, 4 * H, H, device="cuda") * 0.05).to(BF16)
    bias = torch.randn(D * 4 * H, device="cuda") * 0.1
    dh = (torch.randn(S, D * H, device="cuda") * 0.1).to(BF16)
    for _ in range(2):
        gates, h_hist, c_hist, h_last, _, sync = ops.lstm_seq_fwd(x, wih, whh, bias)
        _, sync2 = ops.lstm_bwd(gates, whh, h_hist, c_hist, dh, whole_sequence=True)
    torch.cuda.synchronize()
    print(name, "timeouts", int(sync[-1]), int(sync2[-1]), flush=True)
