"""Times the HBM-bound fused kernels at the workload shape (B=256 videos, N=20 clips, D=768): the 4-graph GAT attention
launch (forward / backward, tensor-core fast path vs generic path) and the 3-pair auxiliary-loss launch pair.
Prints achieved GB/s on the algorithmic bytes against the measured HBM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dualvgr_oracle as orc
import dualvgr_videoqa_b200.ops as ops

BF16 = torch.bfloat16
B, N, D, K = int(os.environ.get("TB", 256)), int(os.environ.get("TN", 20)), 768, 4
Dh = D // K
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


adj = orc.build_adjacency(N).cuda()
gate = torch.rand(B, N, device="cuda")
whs = [torch.randn(B * N, D, device="cuda").to(BF16) for _ in range(4)]
avs = [(torch.randn(K, 2 * Dh + 1, device="cuda") * 0.1) for _ in range(4)]
douts = [torch.randn(B * N, D, device="cuda").to(BF16) for _ in range(4)]
d32 = [torch.randn(B, N, D, device="cuda") * 0.5 for _ in range(4)]
gates = [gate] * 4
res = {}
for mode in ("0", "3"):
    os.environ["DVGR_GAT_FAST"] = mode

    def fwd():
        res["o"] = ops.gat_attn_fwd(whs, gates, avs, adj, B, N, p_att=0.15, p_out=0.15, seed=5, want_f32=True)

    def bwd():
        ops.gat_attn_bwd(whs, gates, avs, res["o"][0], douts, adj, B, N, p_att=0.15, p_out=0.15, seed=5, douts32=d32)

    tf, tb = timeit(fwd), timeit(bwd)
    e = B * N * D
    bytes_f = 4 * (2 * e + 2 * e + 4 * e)                   # read Wh, write out bf16 + fp32 copy
    bytes_b = 4 * (2 * e * 3 + 4 * e + 2 * e)               # read Wh, out, dout (bf16) + dout_f32; write dWh
    print(f"GAT x4 graphs [{'tensor-core' if mode == '3' else 'generic'}] B={B} N={N}: fwd {tf * 1e3:.1f} us = {bytes_f / tf / 1e6:.0f} GB/s "
          f"({bytes_f / tf / 1e6 / peak:.2f} of measured {peak:.0f}) | bwd {tb * 1e3:.1f} us = {bytes_b / tb / 1e6:.0f} GB/s ({bytes_b / tb / 1e6 / peak:.2f})",
          flush=True)
os.environ.pop("DVGR_GAT_FAST", None)

ca, cm, aq, mq = (torch.randn(B, N, D, device="cuda") for _ in range(4))
d_ca, d_cm, d_aq, d_mq = (torch.zeros_like(ca) for _ in range(4))
jobs = [dict(x=ca, y=cm, mode=0, coef=1e-3, dx=d_ca, dy=d_cm, acc_x=2, acc_y=2),
        dict(x=aq, y=ca, mode=1, coef=1e-8, dx=d_aq, dy=d_ca, acc_x=0, acc_y=2),
        dict(x=mq, y=cm, mode=1, coef=1e-8, dx=d_mq, dy=d_cm, acc_x=0, acc_y=2)]


def pair():
    ops.pair_loss_multi(jobs, B, N, D, ca)


def aux():
    ops.aux_loss_unit(ca, cm, aq, mq, 1e-3, 1e-8)


tp = timeit(pair)
ta = timeit(aux)
print(f"aux losses, tensor-centric kernels: {ta * 1e3:.1f} us", flush=True)
e4 = B * N * D * 4
print(f"aux losses (3 pairs, value + 4 gradients): {tp * 1e3:.1f} us ; minimal traffic {8 * e4 / 1e6:.0f} MB -> {8 * e4 / tp / 1e6:.0f} GB/s "
      f"({8 * e4 / tp / 1e6 / peak:.2f} of measured)", flush=True)
