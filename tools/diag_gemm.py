"""GPU diagnostic for the tcgen05 GEMM family: each case prints max-abs error vs an fp32 torch product of the same
bf16 inputs. Run one case per process (tools/run_diag.sh) so that a hang cannot take the others down."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dualvgr_videoqa_b200.ops as ops

dev = "cuda"
torch.manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


def report(name, got, ref):
    got = got.float(); ref = ref.float()
    err = (got - ref).abs().max().item()
    rel = err / (ref.abs().max().item() + 1e-9)
    bad = (~torch.isfinite(got)).sum().item()
    print(f"[{name}] max_abs_err={err:.4e} rel_to_max={rel:.3e} nonfinite={bad} {'OK' if rel < 2e-2 and bad == 0 else 'FAIL'}", flush=True)


def case_tn(M, N, K, bn, **kw):
    x, w = rnd(M, K), rnd(N, K, scale=0.1)
    y = ops.linear_fwd(x, w, out_dtype=torch.float32, bn=bn)
    torch.cuda.synchronize()
    report(f"tn M{M} N{N} K{K} bn{bn}", y, x.float() @ w.float().t())


def case_dgrad(M, N, K, bn):
    dy, w = rnd(M, N), rnd(N, K, scale=0.1)
    dx = ops.linear_dgrad(dy, w, bn=bn)
    torch.cuda.synchronize()
    report(f"dgrad M{M} N{N} K{K} bn{bn}", dx, dy.float() @ w.float())


def case_wgrad(M, N, K, bn):
    dy, x = rnd(M, N, scale=0.1), rnd(M, K)
    dw = ops.linear_wgrad(dy, x, bn=bn)
    torch.cuda.synchronize()
    report(f"wgrad M{M} N{N} K{K} bn{bn}", dw, dy.float().t() @ x.float())


def case_epi():
    M, N, K = 300, 200, 136
    x, w = rnd(M, K), rnd(N, K, scale=0.1)
    b = torch.randn(N, device=dev)
    ref = x.float() @ w.float().t() + b
    for act, f in (("none", lambda t: t), ("elu", torch.nn.functional.elu), ("tanh", torch.tanh)):
        y = ops.linear_fwd(x, w, bias=b, act=act)
        torch.cuda.synchronize()
        report(f"epi bias+{act} bf16", y, f(ref))
    c0 = torch.randn(M, N, device=dev)
    c = c0.clone()
    ops.gemm(x, 0, w, 0, M, N, K, c, bias=b, beta=True)
    torch.cuda.synchronize()
    report("epi beta f32", c, ref + c0)
    cb0 = rnd(M, N)
    cb = cb0.clone()
    ops.gemm(x, 0, w, 0, M, N, K, cb, beta=True)
    torch.cuda.synchronize()
    report("epi beta bf16", cb, ref - b + cb0.float())
    # unaligned N (scalar store path) : N = 4002-like
    N2 = 170
    w2 = rnd(N2, K, scale=0.1)
    y2 = torch.zeros(M, N2, device=dev)
    ops.gemm(x, 0, w2, 0, M, N2, K, y2)
    torch.cuda.synchronize()
    report("epi ragged N f32", y2, x.float() @ w2.float().t())
    # row map
    perm = torch.randperm(M, device=dev).int()
    y3 = torch.zeros(M, N, device=dev)
    ops.gemm(x, 0, w, 0, M, N, K, y3, row_map=perm)
    torch.cuda.synchronize()
    ref3 = torch.zeros(M, N, device=dev)
    ref3[perm.long()] = x.float() @ w.float().t()
    report("epi row_map", y3, ref3)


def case_batch():
    # batch 2 with per-batch offsets: A = [2][M][K], B = [2][N][K]
    M, N, K = 200, 256, 192
    x, w = rnd(2, M, K), rnd(2, N, K, scale=0.1)
    y = torch.empty(2, M, N, device=dev)
    ops.gemm(x, 0, w, 0, M, N, K, y, batch=2, c_batch=M * N, a_c2=[0, 1], b_c2=[0, 1])
    torch.cuda.synchronize()
    report("batch2 tn", y, torch.einsum("bmk,bnk->bmn", x.float(), w.float()))
    # column offset on A (c0) : A = [M][2K], use second half
    xx = rnd(M, 2 * K)
    y2 = torch.empty(M, N, device=dev)
    ops.gemm(xx, 0, w[0], 0, M, N, K, y2, a_c0=[K])
    torch.cuda.synchronize()
    report("a_c0 offset", y2, xx[:, K:].float() @ w[0].float().t())
    # segmented MN-major reduction: dW = sum_t dY[t]^T X[t]; A=[T][S][N], B=[T'][S][K] with reversed slot order
    T, S, N, K = 3, 100, 128, 192
    dy, x = rnd(T, S, N, scale=0.1), rnd(T, S, K)
    dw = torch.empty(N, K, device=dev)
    kin = (S + 63) // 64
    ops.gemm(dy, 1, x, 1, N, K, T * kin * 64, dw, k_inner=kin, a_c2=[T - 1], a_c2_step=[-1], b_c2=[0], b_c2_step=[1])
    torch.cuda.synchronize()
    ref = sum(dy[T - 1 - t].float().t() @ x[t].float() for t in range(T))
    report("segmented wgrad", dw, ref)


def case_perf():
    for (M, N, K, bn) in ((81920, 3072, 2048, 256), (81920, 3072, 2048, 128), (5120, 768, 2048, 256), (5120, 1536, 768, 256), (10240, 768, 768, 128)):
        x, w = rnd(M, K), rnd(N, K, scale=0.05)
        b = torch.randn(N, device=dev)
        y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            ops.gemm(x, 0, w, 0, M, N, K, y, bias=b, bn=bn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            ops.gemm(x, 0, w, 0, M, N, K, y, bias=b, bn=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"[perf tn M{M} N{N} K{K} bn{bn}] {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        # sampled check
        idx = torch.randint(0, M, (64,), device=dev)
        report("  sampled", y[idx], x[idx].float() @ w.float().t() + b)
        for _ in range(2):
            t0 = time.time(); yy = x @ w.t(); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            yy = torch.nn.functional.linear(x, w)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"   cuBLAS same shape: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    # wgrad big
    M, N, K = 81920, 3072, 2048
    dy, x = rnd(M, N, scale=0.05), rnd(M, K)
    dw = torch.empty(N, K, device=dev)
    for bn in (256, 128):
        for _ in range(2):
            ops.linear_wgrad(dy, x, out=dw, bn=bn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.linear_wgrad(dy, x, out=dw, bn=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"[perf wgrad M{M} N{N} K{K} bn{bn}] {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    ref = dy[:, :64].float().t() @ x.float()
    report("  wgrad big rows0-63", dw[:64], ref)


def lstm_ref(gx, whh, T, S, H, D):
    """fp32 torch reference of the interleaved-gate LSTM (autograd-capable). gx [T,S,D*4H], whh [D,4H,H]."""
    hs = []
    gates_all = []
    for d in range(D):
        h = torch.zeros(S, H, device=dev); c = torch.zeros(S, H, device=dev)
        for s in range(T):
            t = s if d == 0 else T - 1 - s
            pre = gx[t, :, d * 4 * H:(d + 1) * 4 * H] + h @ whh[d].t()
            pre = pre.view(S, H, 4)
            i, f, g, o = torch.sigmoid(pre[..., 0]), torch.sigmoid(pre[..., 1]), torch.tanh(pre[..., 2]), torch.sigmoid(pre[..., 3])
            c = f * c + i * g
            h = o * torch.tanh(c)
        hs.append(h)
    return torch.cat(hs, dim=1)


def case_lstm():
    for (T, S, H, D) in ((3, 128, 64, 1), (4, 200, 128, 2), (16, 640, 384, 2)):
        gx = (torch.randn(T, S, D * 4 * H, device=dev)).to(torch.bfloat16)
        whh = (torch.randn(D, 4 * H, H, device=dev) * 0.08).to(torch.bfloat16)
        gx32 = gx.float().requires_grad_(True)
        whh32 = whh.float().requires_grad_(True)
        ref = lstm_ref(gx32, whh32, T, S, H, D)
        g = gx.clone()
        h_hist, c_hist, h_last, _ = ops.lstm_fwd(g, whh)
        torch.cuda.synchronize()
        report(f"lstm fwd T{T} S{S} H{H} D{D}", h_last, ref)
        dh = (torch.randn(S, D * H, device=dev)).to(torch.bfloat16)
        ref.backward(dh.float())
        ops.lstm_bwd(g, whh, h_hist, c_hist, dh)
        torch.cuda.synchronize()
        report(f"lstm bwd dgates T{T} S{S} H{H} D{D}", g, gx32.grad)
        # timing of the recurrence
        if S >= 640:
            S2 = 5120
            gx = torch.randn(T, S2, D * 4 * H, device=dev).to(torch.bfloat16)
            for _ in range(2):
                g = gx.clone(); hh, cc, hl, _ = ops.lstm_fwd(g, whh)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g = gx.clone()
            e0.record(); hh, cc, hl, _ = ops.lstm_fwd(g, whh); e1.record(); torch.cuda.synchronize()
            print(f"[perf lstm fwd S{S2} T{T}] {e0.elapsed_time(e1):.3f} ms", flush=True)
            dh = torch.randn(S2, D * H, device=dev).to(torch.bfloat16)
            e0.record(); ops.lstm_bwd(g, whh, hh, cc, dh); e1.record(); torch.cuda.synchronize()
            print(f"[perf lstm bwd S{S2} T{T}] {e0.elapsed_time(e1):.3f} ms", flush=True)


def case_sweep():
    """Separates fixed cost, per-tile cost and per-k-block cost of the GEMM kernel."""
    def t(M, N, K, bn, out_f32=False, bias=True, n=20):
        x, w = rnd(M, K), rnd(N, K, scale=0.05)
        b = torch.randn(N, device=dev) if bias else None
        y = torch.empty(M, N, dtype=torch.float32 if out_f32 else torch.bfloat16, device=dev)
        for _ in range(3):
            ops.gemm(x, 0, w, 0, M, N, K, y, bias=b, bn=bn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.gemm(x, 0, w, 0, M, N, K, y, bias=b, bn=bn)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    for bn in (128, 256):
        ncol = bn
        for tiles_per_cta in (1, 2, 4, 8):
            M = 128 * 148 * tiles_per_cta
            row = []
            for K in (64, 256, 768, 2048):
                row.append(f"K{K}:{t(M, ncol, K, bn):7.1f}us")
            print(f"[sweep bn{bn} tiles/CTA={tiles_per_cta}] " + " ".join(row), flush=True)
    print("[sweep nobias f32out bn256 tiles/CTA=4] " + " ".join(f"K{K}:{t(128*148*4, 256, K, 256, True, False):7.1f}us" for K in (64, 768)), flush=True)


def case_epi_prof():
    M = 128 * 148 * 8
    x, w = rnd(M, 64), rnd(256, 64)
    y = torch.empty(M, 256, dtype=torch.bfloat16, device=dev)
    for _ in range(4):
        ops.gemm(x, 0, w, 0, M, 256, 64, y)
        torch.cuda.synchronize()


CASES = {
    "tn_1tile": lambda: case_tn(128, 128, 64, 128),
    "tn_k4": lambda: case_tn(128, 128, 256, 128),
    "tn_256": lambda: case_tn(256, 256, 512, 256),
    "tn_multi": lambda: (case_tn(1024, 768, 768, 128), case_tn(1024, 768, 768, 256), case_tn(4096, 1536, 2048, 256)),
    "tn_ragged": lambda: (case_tn(200, 136, 72, 128), case_tn(333, 264, 200, 256), case_tn(32, 32, 1536, 128)),
    "dgrad": lambda: (case_dgrad(128, 128, 128, 128), case_dgrad(512, 768, 768, 128), case_dgrad(300, 200, 264, 256), case_dgrad(1024, 1536, 768, 256)),
    "wgrad": lambda: (case_wgrad(128, 128, 128, 128), case_wgrad(1024, 768, 768, 128), case_wgrad(1000, 200, 264, 256), case_wgrad(5120, 768, 2048, 256)),
    "epi": case_epi,
    "batch": case_batch,
    "lstm": case_lstm,
    "sweep": case_sweep,
    "epi_prof": lambda: case_epi_prof(),
    "perf": case_perf,
}

if __name__ == "__main__":
    name = sys.argv[1]
    print(f"=== case {name}", flush=True)
    CASES[name]()
    torch.cuda.synchronize()
    print(f"=== case {name} done", flush=True)
