"""CPU-side checks of the host logic around the kernels (no CUDA needed): the blocked warp-tile layout helpers of the
whole-sequence LSTM kernels against the index formulas the device code uses (csrc/gemm.cuh), the split-K cost model of the
weight-gradient GEMMs, and the train engine's flat-buffer layout (gradient buckets, contiguous operand groups)."""
import random

import torch

import dualvgr_oracle as orc


def test_lstm_blocked_layout_matches_device_index_formulas():
    from dualvgr_videoqa_b200 import ops
    random.seed(0)
    T, S, D, H = 3, 70, 2, 64
    RB, UG = (S + 31) // 32, H // 8
    g = torch.randn(T, S, D * 4 * H).to(torch.bfloat16)
    gb = ops.lstm_block_gates(g, D)
    assert tuple(gb.shape) == (T, D, RB, UG, 4, 32, 8) and torch.equal(ops.lstm_unblock_gates(gb, S), g)
    flat = gb.reshape(-1)
    for _ in range(300):
        t, d, seq, j, gate = (random.randrange(n) for n in (T, D, S, H, 4))
        rb, r, ug, u = seq // 32, seq % 32, j // 8, j % 8
        # lstm_blk_gates(t, d, D, RB, UG, rb, ug) = ((((t*D + d)*RB + rb)*UG + ug) * 1024 ; piece q = u/2 at + q*256 ; row r at + r*8
        off = ((((t * D + d) * RB + rb) * UG + ug) * 1024) + (u // 2) * 256 + r * 8 + (u % 2) * 4 + gate
        assert flat[off] == g[t, seq, d * 4 * H + 4 * j + gate]
    c = torch.randn(D, T + 1, S, H)
    cb = ops.lstm_block_c(c)
    assert tuple(cb.shape) == (D, T + 1, RB, UG, 2, 32, 4) and torch.equal(ops.lstm_unblock_c(cb, S), c)
    flat = cb.reshape(-1)
    for _ in range(300):
        d, sl, seq, j = (random.randrange(n) for n in (D, T + 1, S, H))
        rb, r, ug, u = seq // 32, seq % 32, j // 8, j % 8
        # lstm_blk_c(d, slot, T, RB, UG, rb, ug) = ((((d*(T+1) + slot)*RB + rb)*UG + ug) * 256 ; piece q = u/4 at + q*128 ; row r at + r*4
        off = (((((d * (T + 1) + sl) * RB + rb) * UG + ug) * 256)) + (u // 4) * 128 + r * 4 + (u % 4)
        assert flat[off] == c[d, sl, seq, j]


def test_wgrad_split_cost_model():
    from dualvgr_videoqa_b200 import ops
    # the big appearance W_ih gradient: 192 tiles of 128 x 256 on 148 CTAs -> a 3-way split fills 4 waves almost exactly
    bn, ks = ops.wgrad_split(3072, 2048, 81920)
    assert bn == 256 and ks == 3
    # never more splits than k-blocks / 4, never less than 1, and deterministic
    for rows, cols, red in ((768, 768, 5120), (1536, 384, 81920), (6144, 304, 5120), (128, 64, 100)):
        bn, ks = ops.wgrad_split(rows, cols, red)
        assert bn in (128, 256) and 1 <= ks <= max(1, ((red + 63) // 64) // 4)
        assert (bn, ks) == ops.wgrad_split(rows, cols, red)


def test_engine_flat_layout_buckets_and_groups():
    """TrainEngine on a CPU model: parameters become views of ONE flat buffer, the 'late' bucket holds exactly the three
    input encoders, every operand group the kernels treat as one matrix is contiguous (so a single GEMM writes its gradient
    block), and slices stay 16-byte aligned."""
    import dualvgr_videoqa_b200.model.models as M
    import dualvgr_videoqa_b200.autograd as ag
    from dualvgr_videoqa_b200.engine import TrainEngine
    V, A, U = 30, 10, 2
    model = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=8, graph_module="GAT", graph_layers=1, unit_layers=U)
    model.load_state_dict(orc.make_state_dict(U, A, V), strict=True)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    eng = TrainEngine(model)
    try:
        assert eng.flat.numel() == eng.numel and 0 < eng.late_numel < eng.numel
        late_ids = {id(p) for name in ("visual_appearance_input_unit", "linguistic_input_unit", "visual_motion_input_unit")
                    for p in getattr(model, name).parameters()}
        base = eng.flat.data_ptr()
        for n, p in model.named_parameters():
            assert torch.equal(p.detach(), before[n]), n                      # values preserved
            off = (p.data_ptr() - base) // 4
            assert 0 <= off < eng.numel and off % 4 == 0                       # a view of the flat buffer, 16-byte aligned
            assert (off < eng.late_numel) == (id(p) in late_ids), n            # bucket membership
            assert p.grad is not None and (p.grad.data_ptr() - eng.gflat.data_ptr()) // 4 == off
        for grp in ag.grad_groups(model):
            tgt = ag.grad_target(grp)
            assert tgt is not None and tgt.shape[0] == sum(g.numel() // g.shape[-1] for g in grp)
    finally:
        eng.close()
        ag.DIRECT_GRAD[0] = False
