"""PyTorch-eager GPU baseline for the DualVGR train step — TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).

SURVEY.md §2.2 / §8d and BASELINE.md §5.2 name "PyTorch eager on B200 running the reference modules" as the number to beat.
The reference itself (/root/reference) cannot travel to the GPU box, so this file restates its eager execution on top of
the oracle's formulas (oracle/dualvgr_oracle.py), in two flavours:

  lean      the oracle's fused formulas with cuDNN LSTMs (nn.LSTM, as the reference's encoders, model/Preprocessing.py:97-101,
            202) — FASTER than the reference can be: no [B,N,N,2Dh] pair tensor, no .cpu() round trips, no per-sample loops
  faithful  adds the reference's characteristic eager costs back: the materialised pair tensor of
            _prepare_attentional_mechanism_input (model/GraphNN.py:115-155), the per-sample mask loop of QueryAttn
            (model/utils.py:72-75), the .cpu() / .cuda() round trip of the 4U GAT outputs (model/models.py:153-160,
            train.py:152-153) and the per-sample torch.trace loop of loss_dependence (utils.py:28-31)

Both run the loop body of train.py:139-159 (forward, CE + alpha common + beta HSIC, backward, clip 12, Adam), dropout on.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import dualvgr_oracle as orc


class EagerDualVGR:
    def __init__(self, U, A, V, N, device, faithful=False, dtype=torch.float32):
        self.U, self.N, self.faithful, self.dev = U, N, faithful, device
        sd = orc.make_state_dict(U, A, V)
        self.P = {}
        for k, v in sd.items():
            v = v.to(device)
            self.P[k] = nn.Parameter(v.to(dtype)) if (v.is_floating_point() and "running_" not in k) else v
        self.adj = orc.build_adjacency(N).to(device=device, dtype=dtype)
        H, W, Dv = 384, 300, 2048
        self.lstm = {}
        for prefix, inp in (("linguistic_input_unit.concatRNN.rnn", W), ("linguistic_input_unit.encoder", W),
                            ("visual_appearance_input_unit.encoder", Dv)):
            m = nn.LSTM(inp, H, batch_first=True, bidirectional=True).to(device=device, dtype=dtype)
            for name in list(m._parameters):
                m._parameters[name] = self.P[f"{prefix}.{name}"]
            m._init_flat_weights()
            m.flatten_parameters()
            self.lstm[prefix] = m
        self.params = [p for p in self.P.values() if isinstance(p, nn.Parameter)]
        self.opt = torch.optim.Adam(self.params, lr=1e-4)

    # ---- encoders on cuDNN (reference: nn.LSTM everywhere)
    def _bilstm(self, prefix, x, lengths=None):
        m = self.lstm[prefix]
        if lengths is None:
            out, (h, _) = m(x)
            return out, torch.cat([h[0], h[1]], dim=-1)
        packed = nn.utils.rnn.pack_padded_sequence(x, lengths.cpu(), batch_first=True, enforce_sorted=False)
        out, (h, _) = m(packed)
        out, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True, total_length=x.shape[1])
        return out, torch.cat([h[0], h[1]], dim=-1)

    def _gat_faithful(self, name, i, x, gate, p=0.15):
        """punishGAT with the reference's materialised [B,N,N,2Dh] attention input (model/GraphNN.py:95-155)."""
        P = self.P
        x = F.dropout(x, p, True)
        B, N, _ = x.shape
        outs = []
        for k in range(4):
            pre = f"visual_input_unit.{name}.{i}.attention_{k}"
            Wh = F.linear(x, P[f"{pre}.W.weight"], P[f"{pre}.W.bias"])
            Dh = Wh.shape[-1]
            rep_chunks = Wh.repeat_interleave(N, dim=1)
            rep_alt = Wh.repeat(1, N, 1)
            a_in = torch.cat([rep_chunks, rep_alt], dim=2).view(B, N, N, 2 * Dh)
            e = F.leaky_relu(F.linear(a_in, P[f"{pre}.a.weight"], P[f"{pre}.a.bias"]).squeeze(3), 0.01)
            att = torch.where(self.adj > 0, e, -9e15 * torch.ones_like(e))
            att = F.dropout(F.softmax(att, dim=-1), p, True)
            outs.append(F.elu(torch.matmul(att, Wh * gate[:, :, None])))
        return F.dropout(torch.cat(outs, dim=2), p, True)

    def _gat_lean(self, name, i, x, gate, p=0.15):
        B, N, D = x.shape
        m_in = F.dropout(torch.ones_like(x), p, True)
        m_att = [F.dropout(torch.ones((B, N, N), device=x.device, dtype=x.dtype), p, True) for _ in range(4)]
        m_out = F.dropout(torch.ones_like(x), p, True)
        return orc.punish_gat(self.P, name, i, x, self.adj, gate, m_in, m_att, m_out)

    def forward(self, app, mot, question, qlen):
        P, U = self.P, self.U
        B, N = app.shape[:2]
        pl = "linguistic_input_unit"
        words = torch.tanh(F.dropout(P[f"{pl}.encoder_embed.weight"][question], 0.15, True))
        dynamic_q, _ = self._bilstm(f"{pl}.concatRNN.rnn", words, qlen)
        _, q_emb = self._bilstm(f"{pl}.encoder", words, qlen)
        q_emb = F.dropout(q_emb, 0.18, True)
        x = torch.tanh(F.dropout(app, 0.15, True)).reshape(B * N, app.shape[2], -1)
        _, h = self._bilstm("visual_appearance_input_unit.encoder", x)
        a = F.dropout(h, 0.18, True).reshape(B, N, -1)
        m = F.linear(mot, P["visual_motion_input_unit.weight"], P["visual_motion_input_unit.bias"])
        gat = self._gat_faithful if self.faithful else self._gat_lean
        ca_l, cm_l, aq_l, mq_l = [], [], [], []
        for i in range(U):
            if self.faithful:       # QueryAttn's per-sample mask loop with a host sync per row (model/utils.py:72-75)
                pq = f"visual_input_unit.queryAttn.{i}"
                y = F.normalize(F.linear(dynamic_q, P[f"{pq}.feat_enhance.weight"], P[f"{pq}.feat_enhance.bias"]), p=2, dim=-1)
                alpha = F.softmax(F.linear(y, P[f"{pq}.fc.weight"], P[f"{pq}.fc.bias"]), dim=1)
                mask = torch.zeros(alpha.shape[:2], device=alpha.device, dtype=alpha.dtype)
                for b in range(B):
                    mask[b, :int(qlen[b])] = 1
                alpha = alpha * mask.unsqueeze(2)
                alpha = alpha / (alpha.sum(1, keepdim=True) + 1e-5)
                q_c = torch.bmm(alpha.transpose(1, 2), words).squeeze(1)
            else:
                q_c, _ = orc.query_attn(P, i, words, dynamic_q, qlen)
            g_a = orc.query_punish(P, "queryPunish_appear", i, q_c, a)
            g_m = orc.query_punish(P, "queryPunish_motion", i, q_c, m)
            com_app, aq = gat("acGCN", i, a, g_a), gat("appearance_GCN", i, a, g_a)
            com_mot, mq = gat("mcGCN", i, m, g_m), gat("motion_GCN", i, m, g_m)
            if self.faithful:       # model/models.py:153-160
                ca_l.append(com_app.cpu()); aq_l.append(aq.cpu()); cm_l.append(com_mot.cpu()); mq_l.append(mq.cpu())
            else:
                ca_l.append(com_app); aq_l.append(aq); cm_l.append(com_mot); mq_l.append(mq)
            aq_embed, _ = orc.attention_sfgcn(P, "attention_appearance", i, com_app, aq)
            mq_embed, _ = orc.attention_sfgcn(P, "attention_motion", i, com_mot, mq)
            a, m = a + aq_embed, m + mq_embed
        visual = orc.mfb(P, a, m)
        pooled = orc.context_self_attn(P, visual, F.dropout(torch.ones_like(visual), 0.15, True))
        ones = torch.ones((B, 2 * 768), device=a.device, dtype=a.dtype)
        logits = orc.output_unit(P, q_emb, pooled, True, F.dropout(ones, 0.15, True), F.dropout(ones[:, :768], 0.15, True))
        return logits, ca_l, cm_l, aq_l, mq_l

    def _hsic_faithful(self, e1, e2):
        """utils.py:20-31 with its per-sample trace loop."""
        dim = self.N
        R = torch.eye(dim, device=self.dev, dtype=e1.dtype) - (1.0 / dim) * torch.ones(dim, dim, device=self.dev, dtype=e1.dtype)
        RK = torch.bmm(torch.matmul(R, torch.bmm(e1, e1.transpose(1, 2))), torch.matmul(R, torch.bmm(e2, e2.transpose(1, 2))))
        out = 0
        for b in range(RK.shape[0]):
            out = out + torch.trace(RK[b])
        return out

    def train_step(self, app, mot, question, qlen, answers, alpha=1.0, beta=1e-8):
        self.opt.zero_grad(set_to_none=True)
        logits, ca, cm, aq, mq = self.forward(app, mot, question, qlen)
        loss = F.cross_entropy(logits.float(), answers)
        com = dep = 0
        for i in range(self.U):
            c1, c2, a1, m1 = ca[i], cm[i], aq[i], mq[i]
            if self.faithful:       # train.py:152-153
                c1, c2, a1, m1 = c1.cuda(), c2.cuda(), a1.cuda(), m1.cuda()
                dep = dep + self._hsic_faithful(a1.float(), c1.float()) + self._hsic_faithful(m1.float(), c2.float())
            else:
                dep = dep + orc.loss_dependence(a1.float(), c1.float(), self.N) + orc.loss_dependence(m1.float(), c2.float(), self.N)
            com = com + orc.common_loss(c1.float(), c2.float())
        loss = loss + alpha * com / self.U + beta * dep / self.U
        loss.backward()
        nn.utils.clip_grad_norm_(self.params, max_norm=12)
        self.opt.step()
        return loss.detach()


def time_eager_step(cfg, batch, steps=3, warmup=2, faithful=False, autocast=False):
    """Median ms per full train step of the eager baseline on the current CUDA device; batch = (app, mot, q, qlen, ans)
    device tensors. Returns (ms, samples/s) or raises (e.g. out of memory) — the caller records the reason."""
    dev = batch[0].device
    model = EagerDualVGR(cfg["U"], cfg["A"], cfg["V"], cfg["N"], dev, faithful=faithful)
    times = []
    for it in range(warmup + steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            model.train_step(*batch)
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            times.append(e0.elapsed_time(e1))
    times.sort()
    ms = times[len(times) // 2]
    del model
    torch.cuda.empty_cache()
    return ms, batch[0].shape[0] / (ms * 1e-3)
