"""Train-step engine on the GPU: flat-buffer clip+Adam against torch.optim.Adam + clip_grad_norm_ on the same gradients,
and whole-step CUDA-graph replay against eager steps."""
import copy

import pytest
import torch

import dualvgr_oracle as orc

pytestmark = pytest.mark.gpu


def no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float):
            m.dropout = 0.0


def make(cfg):
    import dualvgr_videoqa_b200.model.models as M
    B, N, L, A, V, U = cfg
    model = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
    model.load_state_dict(orc.make_state_dict(U, A, V), strict=True)
    no_dropout(model)
    return model.cuda().train(), [t.cuda() for t in orc.make_inputs(B, N, L, A, V)]


def test_flat_optimizer_matches_torch_adam_with_clipping():
    """dvgr_sumsq + dvgr_adam_step on the flat buffers == clip_grad_norm_ + torch.optim.Adam (train.py:85,158-159) on the SAME
    gradients: every step the engine's gradient (flat buffer views) is handed to a torch optimizer on a copy of the model;
    the two parameter sets must stay together to fp32 rounding. (How good that gradient is, is the parity tests' business.)"""
    from dualvgr_videoqa_b200.engine import TrainEngine
    cfg = (4, 8, 6, 10, 30, 1)
    model, batch = make(cfg)
    ref_params = [p.detach().clone().requires_grad_(True) for p in model.parameters()]
    p0 = [p.detach().clone() for p in ref_params]
    eng = TrainEngine(model, lr=1e-3, max_norm=0.05)      # max_norm tiny so the clip is active
    opt = torch.optim.Adam(ref_params, lr=1e-3)
    for step in range(4):
        eng._last_BN = (cfg[0], cfg[1])
        eng.forward_backward(*batch)
        for rp, p in zip(ref_params, model.parameters()):
            rp.grad = p.grad.detach().clone()
        norm = torch.nn.utils.clip_grad_norm_(ref_params, max_norm=0.05)
        assert float(norm) > 0.05                          # the clip really is active
        opt.step()
        eng.optimizer_step()
        num = den = 0.0
        for p1, p2, q in zip(model.parameters(), ref_params, p0):
            num += float((p1.detach() - p2.detach()).double().pow(2).sum()); den += float((p2.detach() - q).double().pow(2).sum())
        assert (num / den) ** 0.5 < 1e-4, (step, (num / den) ** 0.5)
        # the bf16 operand shadow follows the fp32 master weights
        assert torch.equal(eng.shadow.float(), eng.flat.to(torch.bfloat16).float())
    eng.close()


def test_graph_replay_matches_eager_steps():
    from dualvgr_videoqa_b200.engine import TrainEngine
    cfg = (6, 20, 8, 32, 60, 2)
    m1, batch = make(cfg)
    m2 = copy.deepcopy(m1)
    start = torch.cat([p.detach().reshape(-1) for p in m1.parameters()]).clone()
    e1, e2 = TrainEngine(m1, lr=1e-5), TrainEngine(m2, lr=1e-5)
    e1.capture(*batch, warmup=3)
    l1 = [float(e1.replay()) for _ in range(2)]
    l2 = [float(e2.train_step(*batch)) for _ in range(5)][3:]
    # the split-K weight-gradient GEMMs add their partial tiles with fp32 atomics, so two runs of the SAME step sequence are
    # not bit-reproducible; Adam's normalised step turns that into sign flips on near-zero gradient components, and by the
    # fifth step the losses of two identical eager runs already differ by up to 3e-3 (measured). 1e-2 separates that noise
    # from a graph that replays stale inputs or skips work (those show up as 10 % and more).
    assert all(abs(a - b) < 1e-2 * abs(b) for a, b in zip(l1, l2)), (l1, l2)
    c1 = torch.cat([p.detach().reshape(-1) for p in m1.parameters()])
    c2 = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    d = float((c1 - c2).norm() / (c2 - start).norm())
    # relative error of the accumulated update after 5 steps: split-K fp32 atomics make the summation order differ between
    # runs, and Adam's normalised step amplifies that for near-zero gradients
    assert d < 1e-1, d       # (observed 1e-2 .. 5e-2 from run to run: the split-K atomic order is not reproducible)
    # a replay on a different batch really uses the new inputs
    app2 = batch[0] * 0.5
    e1.load_batch(app2, *batch[1:])
    l_new = float(e1.replay())
    assert abs(l_new - l1[-1]) > 1e-6


def test_graph_replay_draws_fresh_dropout_masks():
    import dualvgr_videoqa_b200.model.models as M
    from dualvgr_videoqa_b200.engine import TrainEngine
    B, N, L, A, V, U = 6, 20, 8, 32, 60, 1
    model = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
    model.load_state_dict(orc.make_state_dict(U, A, V), strict=True)
    model = model.cuda().train()
    batch = [t.cuda() for t in orc.make_inputs(B, N, L, A, V)]
    eng = TrainEngine(model, lr=0.0)                   # lr 0: parameters frozen, only the dropout masks can change the loss
    eng.capture(*batch, warmup=3)
    losses = [float(eng.replay()) for _ in range(4)]
    assert len(set(round(x, 6) for x in losses)) > 1, losses


def test_early_gradient_bucket_is_final_when_the_unit_input_hooks_fire():
    """The overlapped data-parallel all-reduce (engine._EarlyBucketHook) starts reducing the 'early' gradient bucket when the
    gradients of the unit stack's inputs have all been produced. Single-GPU check of that contract: snapshot the bucket at
    that moment and compare with its content at the end of the backward pass — nothing may still be written into it."""
    from dualvgr_videoqa_b200.engine import TrainEngine
    import dualvgr_videoqa_b200.ops as ops
    cfg = (6, 20, 8, 32, 60, 2)
    model, batch = make(cfg)
    eng = TrainEngine(model, lr=1e-5)
    assert 0 < eng.late_numel < eng.numel

    class Snap:
        def __init__(self):
            self.pending, self.snap, self.calls = 0, None, 0

        def arm(self, n):
            self.pending, self.snap = n, None

        def __call__(self, grad):
            self.pending -= 1
            self.calls += 1
            if self.pending == 0:
                ops.flush_wgrads()          # as the real hook does: the queued early-bucket gradients land before the reduction
                self.snap = eng.gflat[eng.late_numel:].clone()
                self.late_at_hook = eng.gflat[:eng.late_numel].clone()
            return None

    snap = Snap()
    model._unit_inputs_grad_hook = snap
    eng._last_BN = (cfg[0], cfg[1])
    eng.forward_backward(*batch)
    torch.cuda.synchronize()
    assert snap.calls == 4 and snap.snap is not None
    assert torch.equal(snap.snap, eng.gflat[eng.late_numel:]), "a gradient of the early bucket was written after the hooks fired"
    assert float(snap.snap.abs().sum()) > 0
    # and the late bucket (the three encoders) really is produced afterwards
    assert not torch.equal(snap.late_at_hook, eng.gflat[:eng.late_numel])
    model._unit_inputs_grad_hook = None
    eng.close()


def _oracle_trajectory(cfg, dtype, steps, lr):
    B, N, L, A, V, U = cfg
    sd = orc.cast_state_dict(orc.make_state_dict(U, A, V), dtype)
    ps = [v.requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    app, mot, q, qlen, ans = orc.make_inputs(B, N, L, A, V)
    opt = torch.optim.Adam(ps, lr=lr)
    out = []
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        ro = orc.dualvgr_forward(sd, U, app.to(dtype), mot.to(dtype), q, qlen, training=True)
        loss = torch.nn.functional.cross_entropy(ro[0], ans)
        com = sum(orc.common_loss(ro[3][i], ro[4][i]) for i in range(U))
        dep = sum(orc.loss_dependence(ro[5][i], ro[3][i], N) + orc.loss_dependence(ro[6][i], ro[4][i], N) for i in range(U))
        loss = loss + 1.0 * com / U + 1e-8 * dep / U
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ps, 12.0)
        opt.step()
        out.append(float(loss.detach()))
    return out


@pytest.mark.parametrize("precision,tol0", [("bf16", 2e-2), ("fp32", 1e-4)])
def test_training_trajectory_follows_the_oracle(precision, tol0):
    """Full train steps (CE + alpha common + beta HSIC, clip 12, Adam lr 1e-4, dropout off, train-mode BatchNorm) of the engine
    against the oracle's loop body (train.py:139-159) on the same batch and initial weights. The first loss must match at the
    mode's tolerance. After ONE Adam step the comparison is already ill-posed — the first update is lr * sign(g), so every
    near-zero gradient whose sign is rounding noise moves its weight by the full 1e-4: the oracle's own float32 and float64
    runs are 1 % apart after one step and 2.5x apart after three — so step 1 is held to 3x that float32-vs-float64 gap (floor:
    8x the mode's tolerance: bf16 gradients flip more of those signs), and over eight steps the loss has to fall the way the oracle's does."""
    from dualvgr_videoqa_b200.engine import TrainEngine
    cfg = (8, 8, 6, 10, 30, 2)
    model, batch = make(cfg)
    model.set_precision(precision)
    try:
        eng = TrainEngine(model, lr=1e-4)
        ours = [float(eng.train_step(*batch)) for _ in range(8)]
        eng.close()
    finally:
        model.set_precision("bf16")
    ref64 = _oracle_trajectory(cfg, torch.float64, 8, 1e-4)
    ref32 = _oracle_trajectory(cfg, torch.float32, 2, 1e-4)
    print(f"{precision}: ours {[round(x, 4) for x in ours]}\n      oracle f64 {[round(x, 4) for x in ref64]} f32 {[round(x, 4) for x in ref32]}")
    assert abs(ours[0] - ref64[0]) < tol0 * abs(ref64[0]), (ours[0], ref64[0])
    gap = abs(ref32[1] - ref64[1])
    assert abs(ours[1] - ref64[1]) < max(3 * gap, 8 * tol0 * abs(ref64[1])), (ours[1], ref64[1], ref32[1])
    assert min(ours) < 0.5 * ours[0] and min(ref64) < 0.5 * ref64[0]
