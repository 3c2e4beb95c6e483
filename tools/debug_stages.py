"""Stage-by-stage comparison of the CUDA model against the CPU oracle (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import dualvgr_oracle as orc
import dualvgr_videoqa_b200.model.models as M
from dualvgr_videoqa_b200 import autograd as ag
from dualvgr_videoqa_b200.model.GraphNN import fused_gat_layer
from dualvgr_videoqa_b200.model.utils import pad_last

B, N, L, A, V, U = [int(x) for x in sys.argv[1:7]] if len(sys.argv) > 6 else (3, 20, 9, 32, 50, 3)
training = False
BF16 = torch.bfloat16


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


model = M.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
sd = orc.make_state_dict(U, A, V)
model.load_state_dict(sd, strict=True)
for m in model.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
    if hasattr(m, "dropout") and isinstance(getattr(m, "dropout"), float): m.dropout = 0.0
model = model.cuda().train(training)
app, mot, q, qlen, ans = orc.make_inputs(B, N, L, A, V)
sd64 = orc.cast_state_dict(sd, torch.float64)
with torch.no_grad():
    ag.begin_forward()
    qe, words, dq = model.linguistic_input_unit(q.cuda(), qlen.cuda())
    rqe, rwords, rdq = orc.question_encoder(sd64, q, qlen)
    print("q_emb", rel(qe, rqe), "words", rel(words, rwords), "dq", rel(dq, rdq))
    a = model.visual_appearance_input_unit(app.cuda())
    ra = orc.appearance_encoder(sd64, app.double())
    print("app enc", rel(a, ra))
    mi = ag.ops.prep_features(mot.cuda().view(B * N, -1), 1, False, False)
    m_ = ag.linear(mi, model.visual_motion_input_unit.weight, model.visual_motion_input_unit.bias).view(B, N, -1)
    rm = orc.linear(mot.double(), sd64["visual_motion_input_unit.weight"], sd64["visual_motion_input_unit.bias"])
    print("motion", rel(m_, rm))
    unit = model.visual_input_unit
    adj = unit.appearance_adj
    radj = orc.build_adjacency(N).double()
    xa, xm = a.to(BF16), m_.to(BF16)
    rxa, rxm = ra, rm
    wordsp = pad_last(words).to(BF16); dqb = dq.to(BF16); ql = qlen.cuda().to(torch.int32)
    for i in range(U):
        qc, al = unit.queryAttn[i](wordsp, dqb, ql, word_dim=300)
        rqc, ral = orc.query_attn(sd64, i, rwords, rdq, qlen)
        print(f"L{i} qc", rel(qc[:, :300], rqc), "alpha", rel(al, ral))
        query = torch.cat([unit.queryPunish_appear[i].query(qc), unit.queryPunish_motion[i].query(qc)], dim=1)
        ga, gm = ag.GateFn.apply(xa, xm, query)
        rga = orc.query_punish(sd64, "queryPunish_appear", i, rqc, rxa)
        rgm = orc.query_punish(sd64, "queryPunish_motion", i, rqc, rxm)
        print(f"L{i} gate_a", rel(ga, rga), "gate_m", rel(gm, rgm))
        (za, zm), (ca, aq, cm, mq) = fused_gat_layer([unit.acGCN[i], unit.appearance_GCN[i], unit.mcGCN[i], unit.motion_GCN[i]],
                                                     [0, 0, 1, 1], [xa, xm], [ga, gm], adj, training)
        rca = orc.punish_gat(sd64, "acGCN", i, rxa, radj, rga)
        raq = orc.punish_gat(sd64, "appearance_GCN", i, rxa, radj, rga)
        rcm = orc.punish_gat(sd64, "mcGCN", i, rxm, radj, rgm)
        rmq = orc.punish_gat(sd64, "motion_GCN", i, rxm, radj, rgm)
        print(f"L{i} com_app", rel(ca, rca), "aq", rel(aq, raq), "com_mot", rel(cm, rcm), "mq", rel(mq, rmq))
        xa, ea = unit.attention_appearance[i].fused(za, xa)
        xm, em = unit.attention_motion[i].fused(zm, xm)
        rea, _ = orc.attention_sfgcn(sd64, "attention_appearance", i, rca, raq)
        rem, _ = orc.attention_sfgcn(sd64, "attention_motion", i, rcm, rmq)
        rxa, rxm = rxa + rea, rxm + rem
        print(f"L{i} embed_a", rel(ea, rea), "embed_m", rel(em, rem), "xa", rel(xa, rxa), "xm", rel(xm, rxm))
    vis = unit.visualfusion([xa, xm])
    rvis = orc.mfb(sd64, rxa, rxm)
    print("mfb", rel(vis, rvis))
    pooled = model.feature_aggregation(vis)
    rpooled = orc.context_self_attn(sd64, rvis)
    print("pooled", rel(pooled, rpooled))
    logits = model.output_unit(qe, pooled)
    rlogits = orc.output_unit(sd64, rqe, rpooled, training)
    print("logits", rel(logits, rlogits))
