"""Functional layer over the C ABI: every function takes torch CUDA tensors, allocates outputs with torch (device
memory and streams are torch's), and launches the library's kernels on the current stream.

No arithmetic of the hot path happens in PyTorch here; there is no fallback if the library is missing."""
import ctypes

import torch

from . import _lib

ACT_NONE, ACT_ELU, ACT_TANH = 0, 1, 2
_ACT = {None: 0, "none": 0, "elu": 1, "tanh": 2}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.DvgrError("dualvgr_b200 ops need CUDA tensors: there is no CPU path")


def operand(t, major):
    """Describe a tensor whose LAST dim is contiguous as a GEMM operand. major 0: K-major (rows x K), 1: MN-major (K x rows).
    Leading dims (up to 2) become outer TMA dims 2, 3."""
    assert t.dtype == torch.bfloat16 and t.stride(-1) == 1 and 2 <= t.dim() <= 4
    op = _lib.Operand()
    op.ptr = t.data_ptr()
    op.major = major
    op.ndim = t.dim()
    shape = list(t.shape)[::-1]
    strides = list(t.stride())[::-1]
    for i in range(t.dim()):
        op.dims[i] = shape[i]
        op.strides[i] = strides[i]
    return op


def gemm(A, a_major, B, b_major, M, N, K, C, *, ldc=None, bias=None, act=None, beta=False, batch=1, c_batch=0,
         bias_batch=0, row_map=None, a_c0=None, a_c2=None, a_c3=None, b_c0=None, b_c2=None, b_c3=None, k_inner=0,
         a_c2_step=None, b_c2_step=None, bn=0, max_ctas=0):
    """Raw tcgen05 GEMM: C[b][m][n] = act(sum_k A*B + bias) (+C)."""
    _check_cuda(A, B, C, bias)
    args = _lib.GemmArgs()
    args.A = operand(A, a_major)
    args.B = operand(B, b_major)
    args.M, args.N, args.K, args.batch = M, N, K, batch
    for name, val in (("a_c0", a_c0), ("a_c2", a_c2), ("a_c3", a_c3), ("b_c0", b_c0), ("b_c2", b_c2), ("b_c3", b_c3),
                      ("a_c2_step", a_c2_step), ("b_c2_step", b_c2_step)):
        if val is not None:
            arr = getattr(args, name)
            for i, v in enumerate(val):
                arr[i] = int(v)
    args.k_inner = k_inner
    args.C = C.data_ptr()
    args.ldc = ldc if ldc is not None else C.stride(-2)
    args.c_batch = c_batch
    assert C.dtype in (torch.float32, torch.bfloat16)
    args.out_f32 = 1 if C.dtype == torch.float32 else 0
    args.act = _ACT[act] if not isinstance(act, int) else act
    args.beta = 1 if beta else 0
    if bias is not None:
        assert bias.dtype == torch.float32
        args.bias = bias.data_ptr()
    args.bias_batch = bias_batch
    if row_map is not None:
        assert row_map.dtype == torch.int32
        args.row_map = row_map.data_ptr()
    args.bn = bn
    args.max_ctas = max_ctas
    _lib.check(_lib.gemm(ctypes.byref(args), _stream()), "dvgr_gemm")
    return C


def linear_fwd(x, w, bias=None, act=None, out=None, out_dtype=torch.bfloat16, bn=0):
    """y[M,N] = act(x[M,K] @ w[N,K]^T + bias). x, w bf16; bias fp32."""
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    return gemm(x, 0, w, 0, M, N, K, out, bias=bias, act=act, bn=bn)


def linear_dgrad(dy, w, out=None, beta=False, bn=0):
    """dx[M,K] = dy[M,N] @ w[N,K]  (w read MN-major: no transposed copy)."""
    M, N = dy.shape
    K = w.shape[1]
    if out is None:
        out = torch.empty((M, K), dtype=torch.bfloat16, device=dy.device)
    return gemm(dy, 0, w, 1, M, K, N, out, beta=beta, bn=bn)


def linear_wgrad(dy, x, out=None, beta=False, row_map=None, bn=0):
    """dw[N,K] (fp32) (+)= dy[M,N]^T @ x[M,K]  (both read MN-major)."""
    M, N = dy.shape
    K = x.shape[1]
    if out is None:
        out = torch.empty((N, K), dtype=torch.float32, device=dy.device)
    return gemm(dy, 1, x, 1, N, K, M, out, beta=beta, row_map=row_map, bn=bn)


def gemm_reference(A, a_rs, a_ks, B, b_rs, b_ks, M, N, K):
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    _lib.check(_lib.gemm_reference(_ptr(A), a_rs, a_ks, _ptr(B), b_rs, b_ks, _ptr(C), N, M, N, K, _stream()),
               "dvgr_gemm_reference")
    return C


# ----------------------------------------------------------------------------------------------------- fused LSTM
def _lstm_args(gates, whh, h_hist, c_hist, S, H, T, ndir):
    a = _lib.LstmArgs()
    a.S, a.H, a.T, a.ndir = S, H, T, ndir
    a.gates, a.whh, a.h_hist, a.c_hist = gates.data_ptr(), whh.data_ptr(), h_hist.data_ptr(), c_hist.data_ptr()
    return a


def lstm_fwd(gates, whh, seq_len=None, want_seq=False):
    """Runs the T recurrent steps of all directions.
    gates [T,S,D*4H] bf16 (x W_ih^T + b, gate-interleaved; overwritten with the activated gates), whh [D,4H,H] bf16.
    Returns (h_hist [D,T+1,S,H] bf16, c_hist [D,T+1,S,H] f32, h_last [S,D*H] bf16, seq_out [S,T,D*H] bf16 | None)."""
    _check_cuda(gates, whh)
    T, S, G = gates.shape
    D, H4, H = whh.shape
    assert G == D * H4 and H4 == 4 * H and gates.is_contiguous() and whh.is_contiguous()
    dev = gates.device
    h_hist = torch.empty((D, T + 1, S, H), dtype=torch.bfloat16, device=dev)
    c_hist = torch.empty((D, T + 1, S, H), dtype=torch.float32, device=dev)
    h_hist[:, 0].zero_()
    c_hist[:, 0].zero_()
    h_last = torch.empty((S, D * H), dtype=torch.bfloat16, device=dev)
    seq_out = torch.empty((S, T, D * H), dtype=torch.bfloat16, device=dev) if want_seq else None
    a = _lstm_args(gates, whh, h_hist, c_hist, S, H, T, D)
    a.h_last, a.h_last_ld = h_last.data_ptr(), D * H
    if seq_len is not None:
        assert seq_len.dtype == torch.int32
        a.seq_len = seq_len.data_ptr()
    if seq_out is not None:
        a.seq_out, a.seq_out_ld = seq_out.data_ptr(), D * H
    st = _stream()
    for s in range(T):
        a.s = s
        _lib.check(_lib.lstm_step_fwd(ctypes.byref(a), st), "dvgr_lstm_step_fwd")
    return h_hist, c_hist, h_last, seq_out


def lstm_bwd(gates, whh, h_hist, c_hist, dh_last, seq_len=None, dh_seq=None):
    """Backward through the T steps; `gates` (activated gates from lstm_fwd) is overwritten in place with the
    pre-activation gate gradients [T,S,D*4H], which feed the W_ih / W_hh / bias wgrads."""
    _check_cuda(gates, whh, dh_last)
    T, S, G = gates.shape
    D, H4, H = whh.shape
    a = _lstm_args(gates, whh, h_hist, c_hist, S, H, T, D)
    dc = torch.zeros((D, S, H), dtype=torch.float32, device=gates.device)
    a.dc = dc.data_ptr()
    if dh_last is not None:
        assert dh_last.dtype == torch.bfloat16 and dh_last.stride(-1) == 1
        a.dh_last, a.dh_last_ld = dh_last.data_ptr(), dh_last.stride(0)
    if seq_len is not None:
        a.seq_len = seq_len.data_ptr()
    if dh_seq is not None:
        a.dh_seq, a.seq_out_ld = dh_seq.data_ptr(), D * H
    st = _stream()
    for s in range(T - 1, -1, -1):
        a.s = s
        _lib.check(_lib.lstm_step_bwd(ctypes.byref(a), st), "dvgr_lstm_step_bwd")
    return gates
