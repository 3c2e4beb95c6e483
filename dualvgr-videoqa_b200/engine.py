"""Data-parallel train step for DualVGR on B200: the loop body of the reference's train.py:131-160 with the device
work restated B200-first.

  * one process per GPU; parameters live in ONE flat fp32 buffer, gradients in another (views handed to autograd), so the
    gradient exchange is a single NCCL all-reduce over NVLink and the optimizer is two kernel launches
  * loss = CE + alpha * mean_l common_loss + beta * mean_l (HSIC_app + HSIC_mot)  (train.py:146-154) from the fused
    cross-entropy / pair-loss kernels (value and gradient in the same launch)
  * clip_grad_norm_(12) + Adam(lr)  (train.py:85,158-159) = dvgr_sumsq + dvgr_adam_step on the flat buffers

Videos are independent, so ranks shard the batch and never exchange activations. Two cross-sample couplings remain and
are handled as standard DDP does: BatchNorm uses per-rank batch statistics; CE / common_loss are means
(gradient averaging reproduces the global-batch gradient), HSIC is a SUM over the batch, so its coefficient is multiplied
by world_size before averaging."""
import weakref

import torch
import torch.distributed as dist

from . import autograd as ag
from . import ops


class _EarlyBucketHook:
    """Tensor hook shared by the unit stack's inputs (model/models.py: app, mot, dynamic_q, words): when the last of their
    gradients has been produced, every parameter gradient of the early bucket is final -> start its all-reduce on a side
    stream (captured into the step's CUDA graph as a fork; TrainEngine.optimizer_step joins it)."""

    def __init__(self, engine):
        self.engine = engine
        self.side = torch.cuda.Stream()
        self.pending = 0
        self.fired = False

    def arm(self, n):
        self.pending = n
        self.fired = False

    def __call__(self, grad):
        self.pending -= 1
        if self.pending == 0:
            e = self.engine
            ops.flush_wgrads()       # the queued (deferred) weight gradients of the early bucket must land before it is reduced
            self.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.side):
                dist.all_reduce(e.gflat[e.late_numel:], op=dist.ReduceOp.SUM, group=e.pg)
            self.fired = True
        return None


class TrainEngine:
    def __init__(self, model, lr=1e-4, max_norm=12.0, alpha=1.0, beta=1e-8, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None):
        self.model = model
        self.lr, self.max_norm, self.alpha, self.beta, self.betas, self.eps = lr, max_norm, alpha, beta, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        # flat layout: the groups the fused kernels see as one matrix (contiguous, in order) and everything else, split in
        # two contiguous buckets by WHEN the gradient is final: "late" = the three input encoders (their backward runs last:
        # LSTM recurrences + the big W_ih weight gradients, ~2 ms), "early" = everything downstream of them. With more than
        # one rank the early bucket is all-reduced on a side stream while the encoders' backward still runs.
        late_ids = set()
        for name in ("visual_appearance_input_unit", "linguistic_input_unit", "visual_motion_input_unit"):
            m = getattr(model, name, None)
            if m is not None:
                late_ids.update(id(p) for p in m.parameters())
        grouped, seen = [], set()
        for grp in ag.grad_groups(model):
            for p in grp:
                if p.requires_grad and id(p) not in seen:
                    grouped.append(p); seen.add(id(p))
        rest = [p for p in model.parameters() if p.requires_grad and id(p) not in seen]
        late = [p for p in grouped if id(p) in late_ids] + [p for p in rest if id(p) in late_ids]
        early = [p for p in grouped if id(p) not in late_ids] + [p for p in rest if id(p) not in late_ids]
        self.params = late + early
        self._n_late = len(late)
        dev = self.params[0].device
        sizes = [(p.numel() + 7) // 8 * 8 for p in self.params]           # 16-byte aligned slices (also in the bf16 shadow)
        total = sum(sizes)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.gflat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(total, dtype=torch.bfloat16, device=dev)   # bf16 GEMM operands, written by the Adam kernel
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.late_numel = sum(sizes[:self._n_late])
        off = 0
        with torch.no_grad():
            for p, n in zip(self.params, sizes):
                sl = self.flat[off:off + p.numel()].view_as(p)
                sl.copy_(p.data)
                p.data = sl
                p.grad = self.gflat[off:off + p.numel()].view_as(p)
                if p.dim() == 2:
                    ag.SHADOW[id(p)] = (weakref.ref(p), self.shadow[off:off + p.numel()].view_as(p))
                off += n
            self.shadow.copy_(self.flat)
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)     # device-side step counter (graph replay)
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)     # device-side dropout seed offset
        self.graph = None
        self.numel = total
        if self.world > 1:      # replicas must start identical (the reference seeds every process the same, train.py:425-428)
            dist.broadcast(self.flat, src=0, group=self.pg)
        ag.invalidate_weight_cache()
        ag.DIRECT_GRAD[0] = True      # weight-gradient GEMMs accumulate straight into the flat gradient buffer
        import os
        overlap_ok = os.environ.get("DVGR_ALLREDUCE_OVERLAP", "1") != "0"        # A/B knob: 0 = one all-reduce after backward
        self._overlap = _EarlyBucketHook(self) if (self.world > 1 and overlap_ok and 0 < self.late_numel < total) else None
        if hasattr(model, "_unit_inputs_grad_hook") or self._overlap is not None:
            model._unit_inputs_grad_hook = self._overlap

    # ------------------------------------------------------------------------------------------------------------
    def loss(self, outputs, answers):
        """(total, ce, loss_com_sum, loss_dep_sum, n_correct) — train.py:146-154 and batch_accuracy (train.py:352-356)."""
        logits, _, _, com_app, com_mot, aq, mq = outputs
        n = len(aq)
        B, N = aq[0].shape[0], aq[0].shape[1]
        ce, correct = ag.CrossEntropyFn.apply(logits, answers)
        total, com, dep = ce, None, None
        c_com, c_dep = (self.alpha / max(n, 1)) / (B * N * N), self.beta * self.world / max(n, 1)
        for i in range(n):
            aux, vals = ag.AuxLossFn.apply(com_app[i], com_mot[i], aq[i], mq[i], c_com, c_dep, True)
            total = total + aux            # coefficient exactly 1: AuxLossFn's unit_grad contract
            c, d = vals[0], vals[1] + vals[2]
            com = c if com is None else com + c
            dep = d if dep is None else dep + d
        if n > 0:       # un-scaled sums, as train.py accumulates them (logging only)
            com, dep = com / (c_com * B * N * N), dep / c_dep if c_dep != 0 else dep
        return total, ce, com, dep, correct

    def train_step(self, app, mot, question, question_len, answers):
        """One optimizer step on this rank's shard. Returns the (device) total loss of the shard."""
        self.model.train()
        self.gflat.zero_()
        outputs = self.model(app, mot, question, question_len)
        total, ce, com, dep, correct = self.loss(outputs, answers)
        ops.DEFER_WGRAD[0] = True          # weight / bias gradients of the nn.Linear layers: queued during backward ...
        try:
            total.backward()
        except BaseException:
            ops.clear_deferred()           # never leave half a step's gradients queued for the next one
            raise
        finally:
            ops.DEFER_WGRAD[0] = False
        ops.flush_wgrads()                 # ... and launched as ONE grouped tcgen05 GEMM + ONE grouped column sum
        self.optimizer_step()
        return total.detach()

    def optimizer_step(self):
        if self.world > 1:
            if self._overlap is not None and self._overlap.fired:
                # the early bucket is already in flight on the side stream (launched from the backward pass); reduce the
                # encoders' bucket here and join
                dist.all_reduce(self.gflat[:self.late_numel], op=dist.ReduceOp.SUM, group=self.pg)
                torch.cuda.current_stream().wait_stream(self._overlap.side)
                self._overlap.fired = False
            else:
                dist.all_reduce(self.gflat, op=dist.ReduceOp.SUM, group=self.pg)
        self.step_count += 1
        self.step_dev += 1
        scale = 1.0 / self.world
        nsq = ops.sumsq(self.gflat)
        ops.adam_step(self.flat, self.gflat, self.m, self.v, self.lr, self.step_count, self.betas[0], self.betas[1],
                      self.eps, max_norm=self.max_norm, norm_sq=nsq, grad_scale=scale, step_dev=self.step_dev,
                      shadow=self.shadow)
        ag.invalidate_weight_cache()

    # ------------------------------------------------------------------------------------------------------------
    # whole-step CUDA graph: forward, losses, backward, all-reduce, clip + Adam, weight re-casts = ONE graph launch.
    # Everything that changes between steps lives on the device (inputs in static buffers, dropout seed offset, Adam step).
    def capture(self, app, mot, question, question_len, answers, warmup=3):
        from . import _lib
        self.static = {k: torch.empty_like(v) for k, v in
                       dict(app=app, mot=mot, q=question, qlen=question_len, ans=answers).items()}
        for k, v in dict(app=app, mot=mot, q=question, qlen=question_len, ans=answers).items():
            self.static[k].copy_(v)
        _lib.lib.dvgr_set_seed_offset(self.seed_dev.data_ptr())

        def body():
            self.seed_dev += 1
            st = self.static
            return self.train_step(st["app"], st["mot"], st["q"], st["qlen"], st["ans"])

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_loss = body()
        self.launches_per_replay = _lib.launch_count() - n0    # library kernels recorded into the graph
        ag.invalidate_weight_cache()
        return self.graph

    def close(self):
        """Detaches the engine's device-side dropout counter from the library (call before dropping the engine)."""
        from . import _lib
        _lib.lib.dvgr_set_seed_offset(None)
        self.graph = None
        for p in self.params:
            ag.SHADOW.pop(id(p), None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_batch(self, app, mot, question, question_len, answers, non_blocking=True):
        st = self.static
        for k, v in dict(app=app, mot=mot, q=question, qlen=question_len, ans=answers).items():
            st[k].copy_(v, non_blocking=non_blocking)

    def replay(self):
        """One captured train step on the batch currently in the static buffers; returns the (device) loss tensor."""
        self.graph.replay()
        self.step_count += 1
        return self.static_loss
