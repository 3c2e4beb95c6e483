"""Whole-subgraph autograd Functions: the DualVGR unit stack (reference model/models.py:121-173) and the question input unit
(model/Preprocessing.py:106-127) each run as ONE torch.autograd.Function whose forward and backward are straight-line
sequences of library launches.

Why: built from one Function per module, a train step spent ~230 of its ~450 kernel launches in framework glue — gradient
accumulation where a tensor feeds several consumers (every unit layer reads the clip features four times), zero fills,
concatenations, slice copies, index arithmetic. Here every such accumulation is folded into the kernel that produces the
gradient (accumulating epilogues, multi-input merges), the appearance and motion streams travel as ONE stacked [2, B*N, D]
tensor so that each per-stream GEMM / attention pair is a single batched launch, and the auxiliary losses of a layer
(train.py:148-154) run on a side stream, overlapped with the rest of the step.

The per-module Functions in autograd.py stay: they back the public forward() of each mirrored nn.Module."""
import os

import torch
from torch.autograd import Function

from . import autograd as ag
from . import ops

BF16, F32 = torch.bfloat16, torch.float32
G = 4                  # graphs per unit layer: acGCN, appearance_GCN (appearance stream), mcGCN, motion_GCN (motion stream)
N_LAYER_PARAMS = 8 + 16 * 4 + 6


def unit_layer_params(unit, i):
    """Parameters of unit layer i in the order UnitStackFn expects (78 tensors)."""
    qa = unit.queryAttn[i]
    qpa, qpm = unit.queryPunish_appear[i].query_weight, unit.queryPunish_motion[i].query_weight
    ps = [qa.feat_enhance.weight, qa.feat_enhance.bias, qa.fc.weight, qa.fc.bias, qpa.weight, qpa.bias, qpm.weight, qpm.bias]
    for gat in (unit.acGCN[i], unit.appearance_GCN[i], unit.mcGCN[i], unit.motion_GCN[i]):
        for att in gat.attentions:
            ps += [att.W.weight, att.W.bias, att.a.weight, att.a.bias]
    for att in (unit.attention_appearance[i], unit.attention_motion[i]):
        ps += [att.project[0].weight, att.project[0].bias, att.project[2].weight]
    return ps


class _LP:
    """Named access to one layer's parameter slice."""

    def __init__(self, ps, heads):
        self.fe_w, self.fe_b, self.fc_w, self.fc_b, self.qa_w, self.qa_b, self.qm_w, self.qm_b = ps[:8]
        gp = ps[8:8 + G * heads * 4]
        self.W = [[gp[(g * heads + k) * 4] for k in range(heads)] for g in range(G)]
        self.Wb = [[gp[(g * heads + k) * 4 + 1] for k in range(heads)] for g in range(G)]
        self.aw = [[gp[(g * heads + k) * 4 + 2] for k in range(heads)] for g in range(G)]
        self.ab = [[gp[(g * heads + k) * 4 + 3] for k in range(heads)] for g in range(G)]
        v = ps[8 + G * heads * 4:]
        self.v_w, self.v_b, self.v_w2 = [v[0], v[3]], [v[1], v[4]], [v[2], v[5]]


def _stack2(a, b, M, D):
    """[2, M, D] view over two [.., D] tensors when b directly follows a in memory (their producers wrote into the halves of
    one buffer), else a stacked copy."""
    if (a.is_contiguous() and b.is_contiguous() and a.dtype == BF16 and b.dtype == BF16
            and b.data_ptr() == a.data_ptr() + M * D * 2
            and a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()):      # halves of ONE allocation
        return torch.as_strided(a, (2, M, D), (M * D, D, 1))
    out = torch.empty((2, M, D), dtype=BF16, device=a.device)
    out[0].copy_(a.reshape(M, D))
    out[1].copy_(b.reshape(M, D))
    return out


_SIDE = {}
LAST_QUERY_CHAIN_BWD = [None]      # CUDA event recorded at the end of the most recent QueryChainFn.backward


def side_stream(device, key="aux"):
    if os.environ.get("DVGR_SIDE_STREAMS", "1") == "0":      # A/B knob: everything on the caller's stream
        return torch.cuda.current_stream()
    k = (str(device), key)
    if k not in _SIDE:
        # the auxiliary losses are filler work (low priority: their CTAs only take SMs the critical path leaves idle); the
        # question encoder is on the critical path of the forward pass (high priority, like the engine's main stream)
        # question stream: one level ABOVE the engine's main stream (-1) — when the question encoder's 48 CTAs retire, the
        # query chain behind it must get those SMs before the appearance encoder's still-pending CTAs do
        _SIDE[k] = torch.cuda.Stream(device=device, priority=0 if key.startswith("aux") else -2)
    _ACTIVE.add(k)
    return _SIDE[k]


_ACTIVE = set()        # side streams handed out since the last join (a stream that took no part in a capture must not be joined)


def join_side_streams(device):
    """The current stream waits for everything queued on the side streams of this device that were used since the last join
    (question encoder, auxiliary losses)."""
    cur = torch.cuda.current_stream()
    for k in sorted(_ACTIVE):
        if k[0] == str(device):
            cur.wait_stream(_SIDE[k])
            _ACTIVE.discard(k)


class _GradSink:
    """Collects parameter gradients of a fused backward. Inside an engine step (parameters' .grad bound to the flat gradient
    buffer) weight gradients and column sums are queued / accumulated in place and the Function returns None for them;
    otherwise they are computed at once and returned to autograd."""

    def __init__(self):
        self.g = {}

    def weight(self, params, dy, x, cols=None):
        """d[params row-concatenated] += dy[M, rows]^T x[M, K]."""
        rows = sum(p.shape[0] for p in params)
        K = params[0].shape[1]
        tgt = ag.grad_target(params)
        if tgt is not None:
            ops.linear_wgrad(dy, x, out=tgt, atomic=True, rows=rows, cols=K)
            return
        dw = ops.linear_wgrad(dy, x)[:rows, :K]
        r = 0
        for p in params:
            self.g[id(p)] = dw[r:r + p.shape[0]]
            r += p.shape[0]

    def colsum(self, params, part):
        """d[params flattened, concatenated] += column sums of part [R, C], C = total numel of params."""
        assert part.shape[1] == sum(p.numel() for p in params), (part.shape, [p.shape for p in params])
        tg = [ag._bias_target(p) for p in params]
        if all(t is not None for t in tg):
            c = 0
            i = 0
            while i < len(params):           # maximal runs of parameters whose .grad views are adjacent: one problem each
                j, n = i + 1, params[i].numel()
                while j < len(params) and tg[j].data_ptr() == tg[i].data_ptr() + n * 4:
                    n += params[j].numel()
                    j += 1
                ops.colsum(part[:, c:c + n], out=torch.as_strided(tg[i], (n,), (1,)), accumulate=True)
                c += n
                i = j
            return
        v = ops.colsum(part if part.stride(1) == 1 else part.contiguous())
        c = 0
        for p in params:
            self.g[id(p)] = v[c:c + p.numel()].view_as(p)
            c += p.numel()

    def grads_for(self, params):
        return tuple(self.g.get(id(p)) for p in params)


def _pack_small(PL, heads, D, Dh, dev):
    """Gathers the small per-layer parameter vectors into the packed fp32 operands of the fused kernels: ONE scatter launch
    per 128 segments for the whole stack (they only change at optimizer steps, but a forward must see the live values)."""
    U = len(PL)
    pk = {
        "avec": torch.empty((U, G, heads, 2 * Dh + 1), dtype=F32, device=dev),
        "gat_b": torch.empty((U, G, D), dtype=F32, device=dev),
        "v_b": torch.empty((U, 2, D), dtype=F32, device=dev),
        "v_w2": torch.empty((U, 2, D), dtype=F32, device=dev),
    }
    segs = []
    for i, p in enumerate(PL):
        for g in range(G):
            for k in range(heads):
                segs.append((pk["avec"][i, g, k, :2 * Dh], p.aw[g][k].detach().reshape(-1)))
                segs.append((pk["avec"][i, g, k, 2 * Dh:], p.ab[g][k].detach().reshape(-1)))
                segs.append((pk["gat_b"][i, g, k * Dh:(k + 1) * Dh], p.Wb[g][k].detach()))
        for s in range(2):
            segs.append((pk["v_b"][i, s], p.v_b[s].detach()))
            segs.append((pk["v_w2"][i, s], p.v_w2[s].detach().reshape(-1)))
    ops.scatter(segs, False)
    return pk


class UnitStackFn(Function):
    """DualVGRUnit_multiple.forward without the final MFB (reference model/models.py:141-169), all U layers.

    cfg = (U, heads, pdrop, W, aux, want_f32): pdrop = GAT dropout rate (0 in eval / parity runs), W = word dimension,
    want_f32 = return the dense fp32 copies of the graph outputs (what the auxiliary losses read; False in inference: bf16
    views are returned instead), aux = None or
    (coef_common, coef_dependence, parts [U, B, 3] f32): with aux the three auxiliary-loss terms of every layer
    (train.py:148-154) and their gradients are computed inside this Function on a side stream (values land in `parts`, the
    gradients are applied in backward with coefficient exactly 1 — the engine's contract, see engine.TrainEngine.loss).
    Inputs: app, mot [B,N,D] bf16; query [U, B, 2D] bf16 (QueryChainFn: the cycle queries of every layer, appearance |
    motion); adj [N,N] f32; then N_LAYER_PARAMS tensors per layer (unit_layer_params; the first 8 of a layer belong to
    QueryChainFn and are not touched here).
    Returns (app_out, mot_out [B,N,D] bf16, aq_embed, mq_embed [B,N,D] bf16, then per layer the dense fp32 outputs of
    acGCN, appearance_GCN, mcGCN, motion_GCN [B,N,D])."""

    @staticmethod
    def forward(ctx, cfg, app, mot, query_all, adj, *params):
        U, heads, pdrop, W, aux, want_f32 = cfg[:6]
        chain_fwd = cfg[6] if len(cfg) > 6 else None        # QueryChainFn's state: per-layer ready events (forward) ...
        ctx.chain = chain_fwd if (chain_fwd is not None and want_f32) else None      # ... and its backward is driven from ours
        B, N, D = app.shape
        M, Dh = B * N, D // heads
        dev = app.device
        query_all = ag._c(query_all)
        PL = [_LP(params[i * N_LAYER_PARAMS:(i + 1) * N_LAYER_PARAMS], heads) for i in range(U)]
        pk = _pack_small(PL, heads, D, Dh, dev)
        X = _stack2(app, mot, M, D)
        seed, sid0 = ag._site(3 * G * U)
        want_f32 = bool(want_f32) or aux is not None      # (grad mode is always off inside Function.forward: the caller decides)
        keep, f32_all, embed = [], [], None
        cur = torch.cuda.current_stream()
        events, aux_jobs = [None] * U, []
        for i, p in enumerate(PL):
            sid = sid0 + 3 * G * i
            # ---- Query Punishment Module: per-clip gates of both streams from this layer's cycle queries (QueryChainFn)
            if chain_fwd is not None and chain_fwd["stream"] != cur:
                cur.wait_event(chain_fwd["events"][i])
            query = query_all[i]
            ga, gm = ops.gate_fwd(X[0].view(B, N, D), X[1].view(B, N, D), query)
            # ---- multi-view GAT: every graph projects its own dropped copy of its stream; ONE batched GEMM, ONE attention launch
            if pdrop > 0:
                xt = torch.empty((G, M, D), dtype=BF16, device=dev)
                ops.dropout_multi([X[0], X[0], X[1], X[1]], [xt[g] for g in range(G)], [sid + g for g in range(G)], pdrop, seed)
                a_c2 = [0, 1, 2, 3]
            else:
                xt, a_c2 = X, [0, 0, 1, 1]
            wb = ag.bf16_rows([w for g in range(G) for w in p.W[g]], tag="gat_all").view(G, D, D)
            wh = torch.empty((G, M, D), dtype=BF16, device=dev)
            ops.gemm(xt, 0, wb, 0, M, D, D, wh, ldc=D, bias=pk["gat_b"][i], batch=G, c_batch=M * D, bias_batch=D, a_c2=a_c2,
                     b_c2=[0, 1, 2, 3])
            z = torch.empty((G, M, D), dtype=BF16, device=dev)          # [stream][common, specific][M][D]
            streams = [sid + G + 2 * g for g in range(G)]
            gates4 = [ga, ga, gm, gm]
            _, o32 = ops.gat_attn_fwd([wh[g] for g in range(G)], gates4, [pk["avec"][i, g] for g in range(G)], adj, B, N,
                                      heads=heads, p_att=pdrop, p_out=pdrop, seed=seed, streams=streams,
                                      outs=[z[g] for g in range(G)], want_f32=want_f32)
            aux_grads = None
            if aux is not None:       # outputs of the auxiliary-loss kernels (launched after the loop, see below)
                aux_grads = tuple(torch.empty_like(t) for t in o32)
                aux_jobs.append((i, o32, aux_grads, ops.aux_loss_workspace(B, N, D, o32[0])))
            if want_f32:
                f32_all += o32
            else:
                f32_all += [z[g].view(B, N, D) for g in range(G)]
            # ---- common-vs-specific view attention + residual, both streams per launch
            w1 = ag.bf16_rows(p.v_w, tag="view").view(2, D, D)
            hidden = torch.empty((G, M, D), dtype=BF16, device=dev)
            ops.gemm(z.view(2, 2 * M, D), 0, w1, 0, 2 * M, D, D, hidden, ldc=D, bias=pk["v_b"][i], act="tanh", batch=2,
                     c_batch=2 * M * D, bias_batch=D, a_c2=[0, 1], b_c2=[0, 1])
            Xn, embed, beta = ops.view_attn_fwd_multi(hidden.view(2, 2, M, D), z.view(2, 2, M, D), X, pk["v_w2"][i],
                                                      want_embed=(i == U - 1))
            keep.append(dict(query=query, ga=ga, gm=gm, xt=xt if pdrop > 0 else None, wh=wh, z=z, hidden=hidden, beta=beta, X=X,
                             wb=wb, w1=w1, aux_grads=aux_grads))
            X = Xn
        if aux_jobs:
            # auxiliary losses of every layer (values + all four gradients each) on the low-priority side stream, forked HERE:
            # they run next to the fusion / read-out / classifier kernels that follow the stack (a few small GEMMs that leave
            # most SMs idle) and the first part of the backward pass; LAST layer first, the order backward needs them in.
            # (Forked right after each layer's graph attention they fought the stack's own GEMMs for SMs: a GEMM CTA needs a
            # whole SM's shared memory and cannot start while any small CTA is resident there — measured +0.4 ms.)
            c_com, c_dep, parts = aux
            n_side = max(1, min(len(aux_jobs), int(os.environ.get("DVGR_AUX_STREAMS", "3"))))
            for r, (i, o32, gr, ws) in enumerate(reversed(aux_jobs)):
                # one low-priority stream per layer (round-robin): the three layers' latency-bound launches overlap each other
                side = side_stream(dev, "aux" if r % n_side == 0 else f"aux{r % n_side}")
                if r < n_side:
                    side.wait_stream(cur)
                with torch.cuda.stream(side):
                    ops.aux_loss_unit_into(o32[0], o32[2], o32[1], o32[3], c_com, c_dep, gr[0], gr[2], gr[1], gr[3], parts[i], ws)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    events[i] = ev
        ctx.set_materialize_grads(False)      # unused outputs (e.g. the fp32 graph outputs in an engine step) arrive as None
        # everything the side-stream kernels touch must outlive them: their workspaces were allocated on THIS stream, and a
        # block freed here could be handed to the next allocation while the (low-priority) side stream still reads it
        ctx.aux_jobs = aux_jobs
        ctx.keep, ctx.pk, ctx.events = keep, pk, events
        ctx.cfg = (U, heads, pdrop, W, B, N, D, seed, sid0)
        ctx.PL, ctx.params = PL, params
        ctx.save_for_backward(adj, query_all)
        if embed is None:
            embed = torch.zeros((2, M, D), dtype=BF16, device=dev)
        outs = (X[0].view(B, N, D), X[1].view(B, N, D), embed[0].view(B, N, D), embed[1].view(B, N, D)) + tuple(f32_all)
        return outs

    @staticmethod
    def backward(ctx, dapp, dmot, dea, dem, *d32_all):
        U, heads, pdrop, W, B, N, D, seed, sid0 = ctx.cfg
        adj, query_all = ctx.saved_tensors
        M, Dh = B * N, D // heads
        dev = adj.device
        dquery_all = torch.empty_like(query_all)
        chain = ctx.chain
        chain_acc = dict(sink=_GradSink(), d_dq=None, dwords=None) if chain is not None else None
        pk, PL = ctx.pk, ctx.PL
        sink = _GradSink()
        cur = torch.cuda.current_stream()

        def pair(a, b):
            if a is None and b is None:
                return None
            if a is None:
                a = torch.zeros_like(b)
            if b is None:
                b = torch.zeros_like(a)
            return _stack2(ag._c(a), ag._c(b), M, D)

        dX = pair(dapp, dmot)
        if dX is None:
            dX = torch.zeros((2, M, D), dtype=BF16, device=dev)
        dembed = pair(dea, dem)
        for i in reversed(range(U)):
            k, p = ctx.keep[i], PL[i]
            sid = sid0 + 3 * G * i
            z, hidden, wh, Xin = k["z"], k["hidden"], k["wh"], k["X"]
            d32 = [None if t is None else ag._c(t) for t in d32_all[G * i:G * i + G]]
            assert all(t is None or t.dtype == F32 for t in d32), "gradients of the fp32 graph outputs must be fp32"
            if k["aux_grads"] is not None:
                cur.wait_event(ctx.events[i])
                if any(t is not None for t in d32):
                    raise RuntimeError("UnitStackFn: auxiliary-loss gradients arrive both from the fused side-stream path and "
                                       "from autograd; use one of the two")
                d32 = list(k["aux_grads"])
            # ---- view attention backward (both streams), then its projection: dz += dhid W1, wgrad / bias queued
            dz, dhid, part = ops.view_attn_bwd_multi(dX, dembed if i == U - 1 else None, hidden.view(2, 2, M, D),
                                                     z.view(2, 2, M, D), pk["v_w2"][i], k["beta"])
            ops.gemm(dhid.view(2, 2 * M, D), 0, k["w1"], 1, 2 * M, D, D, dz.view(2, 2 * M, D), ldc=D, beta=True, batch=2,
                     c_batch=2 * M * D, a_c2=[0, 1], b_c2=[0, 1])
            for s in range(2):
                dh_s = dhid.view(2, 2 * M, D)[s]
                sink.weight([p.v_w[s]], dh_s, z.view(2, 2 * M, D)[s])
                sink.colsum([p.v_b[s]], dh_s)
                sink.colsum([p.v_w2[s]], part[s])
            # ---- graph attention backward (4 graphs), projection dgrad (batched), wgrad / head biases / attention vectors queued
            dz4 = dz.view(G, M, D)
            dwh = torch.empty((G, M, D), dtype=BF16, device=dev)
            streams = [sid + G + 2 * g for g in range(G)]
            gates4 = [k["ga"], k["ga"], k["gm"], k["gm"]]
            _, dgates, dav = ops.gat_attn_bwd([wh[g] for g in range(G)], gates4, [pk["avec"][i, g] for g in range(G)],
                                              [z[g] for g in range(G)], [dz4[g] for g in range(G)], adj, B, N, heads=heads,
                                              p_att=pdrop, p_out=pdrop, seed=seed, streams=streams, douts32=d32,
                                              dwhs=[dwh[g] for g in range(G)], raw=True)
            # ---- gates -> gradient of this layer's cycle queries; the gates' contribution to the layer input goes into dX (the
            #      residual-branch gradient: nobody else reads it any more) so that it only depends on the attention backward
            if chain is not None:
                # gate backward + the question side of this layer's Query Punishment Module run on the question stream NOW,
                # next to the projection dgrad below and under the remaining layers of this loop (small kernels: SMs to
                # spare) — by the time the stack is done only layer 0's chain is left, and the question encoder's own
                # backward starts right behind it
                ev = torch.cuda.Event()
                ev.record(cur)
                with torch.cuda.stream(chain["stream"]):
                    chain["stream"].wait_event(ev)
                    ops.gate_bwd(Xin[0].view(B, N, D), Xin[1].view(B, N, D), k["query"], k["ga"], k["gm"], dgates[0], dgates[1],
                                 dgates[2], dgates[3], dX[0], dX[1], dquery=dquery_all[i])
                    ev_gate = torch.cuda.Event()
                    ev_gate.record()
                    QueryChainFn.layer_backward(chain, i, dquery_all[i], chain_acc)
            else:
                ev_gate = None
                ops.gate_bwd(Xin[0].view(B, N, D), Xin[1].view(B, N, D), k["query"], k["ga"], k["gm"], dgates[0], dgates[1], dgates[2],
                             dgates[3], dX[0], dX[1], dquery=dquery_all[i])
            # ---- projection dgrad (batched over the 4 graphs), then the gradient of the layer input: residual branch (+ gates)
            #      + the (dropped) inputs of the two graphs of each stream
            dxt = torch.empty((G, M, D), dtype=BF16, device=dev)
            ops.gemm(dwh, 0, k["wb"], 1, M, D, D, dxt, ldc=D, batch=G, c_batch=M * D, a_c2=[0, 1, 2, 3], b_c2=[0, 1, 2, 3])
            for g in range(G):
                xg = k["xt"][g] if k["xt"] is not None else Xin[g // 2]
                sink.weight(p.W[g], dwh[g], xg)
                sink.colsum(p.Wb[g], dwh[g])
                for h in range(heads):
                    sink.colsum([p.aw[g][h], p.ab[g][h]], dav[g][:, h * (2 * Dh + 1):(h + 1) * (2 * Dh + 1)])
            if ev_gate is not None:
                cur.wait_event(ev_gate)
            dXin = torch.empty((2, M, D), dtype=BF16, device=dev)
            ops.gat_input_bwd([dxt[g] for g in range(G)], [sid + g for g in range(G)], 2, [dX[0], dX[1]], [dXin[0], dXin[1]],
                              pdrop, seed)
            dX = dXin
        if chain is not None:
            chain["bwd"] = chain_acc
        need = ctx.needs_input_grad
        return ((None, dX[0].view(B, N, D) if need[1] else None, dX[1].view(B, N, D) if need[2] else None,
                 dquery_all if need[3] else None, None) + sink.grads_for(ctx.params))


class QueryChainFn(Function):
    """The QUESTION side of the Query Punishment Module of all U unit layers (reference model/utils.py:60-100, called per layer
    at model/models.py:142-147): feat_enhance GEMM -> word attention -> the two cycle-query projections (one GEMM), giving
    query [U, B, 2D] (appearance | motion). It depends on the question encoder only, so it is hoisted out of the per-layer
    critical path of the unit stack: forward right behind the question encoder on its side stream (under the appearance
    encoder's recurrence), backward on that same stream next to the encoders' backward — the stack's main-stream chain
    loses 4 launches per layer forward and 3 backward.
    cfg = (W, pre | None); dq [B*L, D] bf16 (row stride free), words [B, L, Wp] bf16, qlen [B] int32; then per layer the 8
    tensors feat_enhance.weight/bias, fc.weight/bias, query_weight (appear).weight/bias, query_weight (motion).weight/bias."""

    @staticmethod
    def launch(W, dq, words, qlen, params):
        U = len(params) // 8
        B, L, Wp = words.shape
        D = params[0].shape[0]
        dev = words.device
        P = [params[8 * i:8 * i + 8] for i in range(U)]
        small = {"q_b": torch.empty((U, 2 * D), dtype=F32, device=dev), "fe_b": torch.empty((U, D), dtype=F32, device=dev),
                 "fc_w": torch.empty((U, D), dtype=F32, device=dev), "fc_b": torch.empty((U, 8), dtype=F32, device=dev)}
        segs = []
        for i, (fe_w, fe_b, fc_w, fc_b, qa_w, qa_b, qm_w, qm_b) in enumerate(P):
            segs += [(small["q_b"][i, :D], qa_b.detach()), (small["q_b"][i, D:], qm_b.detach()), (small["fe_b"][i], fe_b.detach()),
                     (small["fc_w"][i], fc_w.detach().reshape(-1)), (small["fc_b"][i, :1], fc_b.detach().reshape(-1))]
        ops.scatter(segs, False)
        words = ag._c(words)
        query = torch.empty((U, B, 2 * D), dtype=BF16, device=dev)
        keep, events = [], []
        for i, (fe_w, fe_b, fc_w, fc_b, qa_w, qa_b, qm_w, qm_b) in enumerate(P):
            we = ag.bf16_rows([fe_w])
            y = ops.linear_fwd(dq, we, bias=small["fe_b"][i])
            qc, alpha, nrm, prob, ssum = ops.qattn_fwd(y.view(B, L, D), small["fc_w"][i], small["fc_b"][i], qlen, words, W, Wp)
            wq = ag.bf16_rows([qa_w, qm_w], out_cols=Wp, tag="cat")
            ops.linear_fwd(qc, wq, bias=small["q_b"][i], out=query[i])
            keep.append(dict(y=y, qc=qc, alpha=alpha, nrm=nrm, prob=prob, ssum=ssum, we=we, wq=wq))
            ev = torch.cuda.Event()
            ev.record()
            events.append(ev)             # layer i's queries are ready: the unit stack waits per layer, not for the whole chain
        return dict(query=query, keep=keep, small=small, words=words, dq=dq, qlen=qlen, params=list(params), events=events,
                    stream=torch.cuda.current_stream(), cfg=(U, B, L, D, W, Wp))

    @staticmethod
    def forward(ctx, cfg, dq, words, qlen, *params):
        W, pre = cfg
        if pre is None:
            pre = QueryChainFn.launch(W, dq, words, qlen, params)
        pre["params"] = list(params)
        ctx.pre = pre
        ctx.params = params
        ctx.stream = torch.cuda.current_stream()
        pre["stream"] = ctx.stream
        return pre["query"]

    @staticmethod
    def layer_backward(pre, i, dquery, acc):
        """Backward of layer i's chain from the gradient of its cycle queries; acc = dict(sink, d_dq, dwords) accumulates over
        the layers. Called either from backward() below or — layer by layer, on the question stream, while the unit stack's
        backward is still running on the main stream — from UnitStackFn.backward (which then leaves the result in
        pre["bwd"] for backward() to hand on)."""
        U, B, L, D, W, Wp = pre["cfg"]
        small, params, k = pre["small"], pre["params"], pre["keep"][i]
        fe_w, fe_b, fc_w, fc_b, qa_w, qa_b, qm_w, qm_b = params[8 * i:8 * i + 8]
        sink = acc["sink"]
        dqc = ops.linear_dgrad(dquery, k["wq"])
        sink.weight([qa_w, qm_w], dquery, k["qc"])
        sink.colsum([qa_b, qm_b], dquery)
        dy, acc["dwords"], dwf_part, dcf_part = ops.qattn_bwd(dqc, k["y"].view(B, L, D), small["fc_w"][i], pre["qlen"], pre["words"],
                                                              W, k["alpha"], k["nrm"], k["prob"], k["ssum"], dwords=acc["dwords"],
                                                              raw=True)
        sink.colsum([fc_w], dwf_part)
        sink.colsum([fc_b], dcf_part)
        dy2 = dy.view(B * L, D)
        acc["d_dq"] = ops.linear_dgrad(dy2, k["we"], out=acc["d_dq"], beta=acc["d_dq"] is not None)
        sink.weight([fe_w], dy2, pre["dq"])
        sink.colsum([fe_b], dy2)
        acc.setdefault("keep", []).append((dquery, dqc, dy))      # side-stream launches: operands outlive the caller's scope

    @staticmethod
    def backward(ctx, dquery_all):
        pre, params = ctx.pre, ctx.params
        U = pre["cfg"][0]
        acc = pre.pop("bwd", None)
        if acc is None:                          # not driven by UnitStackFn.backward: do the whole chain here
            acc = dict(sink=_GradSink(), d_dq=None, dwords=None)
            dquery_all = ag._c(dquery_all)
            for i in reversed(range(U)):
                QueryChainFn.layer_backward(pre, i, dquery_all[i], acc)
        ev = torch.cuda.Event()
        ev.record()
        LAST_QUERY_CHAIN_BWD[0] = ev          # (the data-parallel hook flushes these gradients from another stream)
        need = ctx.needs_input_grad
        return ((None, acc["d_dq"] if need[1] else None, acc["dwords"] if need[2] else None, None)
                + acc["sink"].grads_for(params))


class QuestionInputFn(Function):
    """InputUnitLinguisticDynamic.forward (reference model/Preprocessing.py:106-127) in one Function: embedding lookup +
    dropout + tanh (one launch, written directly as the bf16 operands of the consumers), both BiLSTMs as ONE 4-direction
    length-masked recurrence (directions 0/1 = concatRNN: per-token states, zero rows at padded positions like
    pad_packed_sequence; 2/3 = encoder: final states of the packed run) — no packing, no host sync on question_len.
    cfg = (p_embed,) (0 in eval). tokens [B, L] int64, qlen [B] int32, table [V, W] f32, then the 8 tensors of concatRNN.rnn
    and the 8 of encoder in nn.LSTM order (w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r).
    Returns (dynamic_q [B*L, 2H] bf16 — a column slice of the 4-direction output buffer, row stride 4H —,
             question_embedding [B, 2H] bf16 (column slice, row stride 4H), words [B, L, Wp] bf16 zero-padded)."""

    @staticmethod
    def launch(p_emb, tokens, qlen, table, params):
        """The forward KERNELS (no autograd node yet): returns the state QuestionInputFn.forward adopts. DualVGR.forward calls
        this first (so the launches go out ahead of the appearance encoder's) and applies the Function LAST among the three
        encoders: autograd runs later-created nodes first, so the question encoder's backward — a latency-bound chain on
        ~24 SMs — is enqueued BEFORE the appearance encoder's persistent all-SM launches and overlaps them from the start
        (those claim their tiles dynamically and simply use the SMs that are left). Created first, its backward was queued
        last and — one run in two — only got its SMs after the whole appearance chain: +0.5 ms."""
        B, L = tokens.shape
        W = table.shape[1]
        H = params[1].shape[1]
        Wp = (W + 7) // 8 * 8
        seed, sid = ag._site(1)
        tokens = ag._c(tokens)
        words, x = ops.embed_fwd(tokens, ag._c(table.detach()), Wp, p_emb, seed, sid)
        w_ih = [params[0], params[4], params[8], params[12]]
        w_hh = [params[1], params[5], params[9], params[13]]
        b_ih = [params[4 * d + 2] for d in range(4)]
        b_hh = [params[4 * d + 3] for d in range(4)]
        wih = ag.bf16_rows(w_ih, out_cols=Wp, lstm_H=H, tag="lstm_ih")
        whh = ag.bf16_rows(w_hh, lstm_H=H, tag="lstm_hh").view(4, 4 * H, H)
        bias = ops.lstm_pack_bias([b.detach() for b in b_ih], [b.detach() for b in b_hh], H)
        gates, h_hist, c_hist, h_last, seq_out, sync = ops.lstm_seq_fwd(x, wih, whh, bias, seq_len=qlen, want_seq=True)
        ag.SYNC_WORDS.append(sync)
        return dict(tokens=tokens, words=words, x=x, wih=wih, whh=whh, gates=gates, h_hist=h_hist, c_hist=c_hist,
                    h_last=h_last, seq_out=seq_out, qlen=qlen, cfg=(B, L, W, Wp, H, p_emb, seed, sid))

    @staticmethod
    def forward(ctx, cfg, tokens, qlen, table, *params):
        p_emb, pre = cfg if len(cfg) == 2 else (cfg[0], None)
        if pre is None:
            pre = QuestionInputFn.launch(p_emb, tokens, qlen, table, params)
        B, L, W, Wp, H, p_emb, seed, sid = pre["cfg"]
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(pre["tokens"], pre["words"], pre["x"], pre["wih"], pre["whh"], pre["gates"], pre["h_hist"],
                              pre["c_hist"], qlen)
        ctx.cfg = pre["cfg"]
        ctx.table = table
        ctx.wih_params = [params[0], params[4], params[8], params[12]]
        ctx.whh_params = [params[1], params[5], params[9], params[13]]
        ctx.bias_params = [params[4 * d + j] for d in range(4) for j in (2, 3)]
        dq = pre["seq_out"].view(B * L, 4 * H)[:, :2 * H]
        q = pre["h_last"][:, 2 * H:]
        return dq, q, pre["words"]

    @staticmethod
    def backward(ctx, d_dq, d_q, d_words):
        tokens, words, x, wih, whh, gates, h_hist, c_hist, qlen = ctx.saved_tensors
        B, L, W, Wp, H, p_emb, seed, sid = ctx.cfg
        dev = x.device
        if d_dq is not None:
            d_dq = ag._rows2d(d_dq).view(B, L, -1)
        if d_q is not None:
            d_q = ag._rows2d(d_q)
        if d_dq is None and d_q is None:
            d_q = torch.zeros((B, 2 * H), dtype=BF16, device=dev)
        dh_seq, dh_last = ops.lstm_pack_dh(d_dq, 2, d_q, 2, B, L, 4, H)
        dgates, sync = ops.lstm_bwd(gates, whh, h_hist, c_hist, dh_last, seq_len=qlen, whole_sequence=True,
                                    dh_seq_blocked=dh_seq)
        ag.SYNC_WORDS.append(sync)
        dg = dgates.view(L * B, 16 * H)
        x2 = x.view(L * B, Wp)
        dx = ops.linear_dgrad(dg, wih)                               # [L*B, Wp] time-major gradient of the word vectors
        table = ctx.table
        dtable = None
        if ctx.needs_input_grad[3]:
            tgt = ag._bias_target(table)
            if tgt is None:
                dtable = tgt = torch.zeros_like(table)
            ops.embed_bwd(tokens, words, ag._c(d_words) if d_words is not None else None, dx, W, tgt, p_emb, seed, sid)
        unmap = ag._lstm_unmap(H, 4, dev)
        t_ih, t_hh = ag.grad_target(ctx.wih_params), ag.grad_target(ctx.whh_params)
        if t_ih is not None:
            ops.linear_wgrad(dg, x2, out=t_ih, row_map=unmap, atomic=True, cols=W, dynamic=True)
            dwih = None
        else:
            dwih = ops.linear_wgrad(dg, x2, row_map=unmap)[:, :W]
        dbs = ag._lstm_bias_grads(dg, ctx.bias_params, H)
        kin = (B + 63) // 64
        dwhh = t_hh.view(4, 4 * H, H) if t_hh is not None else torch.empty((4, 4 * H, H), dtype=F32, device=dev)
        # (the step-0 term is skipped: h_0 = 0, and its slot is not initialised in the whole-sequence layout)
        ops.gemm(dgates, 1, h_hist, 1, 4 * H, H, (L - 1) * kin * 64, dwhh, ldc=H, batch=4, c_batch=4 * H * H,
                 row_map=ag._lstm_unmap(H, 1, dev), a_c0=[4 * H * d for d in range(4)], a_c2=[1, L - 2, 1, L - 2],
                 a_c2_step=[1, -1, 1, -1], b_c2=[1, 1, 1, 1], b_c3=[0, 1, 2, 3], b_c2_step=[1, 1, 1, 1], k_inner=kin,
                 beta=2 if t_hh is not None else 0, dynamic=True)
        grads = []
        for d in range(4):
            sl = slice(4 * H * d, 4 * H * (d + 1))
            grads += [None if dwih is None else dwih[sl], None if t_hh is not None else dwhh[d], dbs[2 * d], dbs[2 * d + 1]]
        return (None, None, None, dtable) + tuple(grads)
