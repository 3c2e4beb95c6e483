"""Validation path and input pipeline on the GPU, plus the reference's own loop bodies replayed against the
mirror: train.py:131-160 (forward, CE + auxiliary losses through `from utils import *`, backward, clip_grad_norm_, Adam,
batch_accuracy) and validate.py:44-134 (eval forward, argmax, per-question-type / per-category bookkeeping) — restated here
statement by statement (the reference's files do not travel to the GPU box), driven by a stub cfg, with the mirror imported
the way train.py does (`import model.models as modelset`, `from utils import *`)."""
import types

import numpy as np
import pytest
import torch

import dualvgr_oracle as orc

pytestmark = pytest.mark.gpu


def _build(cfg, dropout=True):
    import dualvgr_videoqa_b200.model.models as modelset
    B, N, L, A, V, U = cfg
    model = modelset.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
    model.load_state_dict(orc.make_state_dict(U, A, V), strict=True)
    return model.cuda()


def test_accuracy_counters_match_the_reference_bookkeeping():
    import dualvgr_videoqa_b200.ops as ops
    g = torch.Generator().manual_seed(1)
    B, A, V, L = 333, 57, 40, 9
    logits = torch.randn((B, A), generator=g).cuda()
    logits[5, 3] = logits[5, 7] = 99.0                                   # a tie: the first index wins (torch.argmax)
    answers = torch.randint(0, A, (B,), generator=g).cuda()
    answers[:100] = logits[:100].argmax(1)
    question = torch.randint(0, V, (B, L), generator=g).cuda()
    words = {"what": 4, "who": 9, "how": 11, "when": 17, "where": 23}
    table = torch.full((V,), -1, dtype=torch.int32)
    for c, w in enumerate(("what", "who", "how", "when", "where")):
        table[words[w]] = c
    counts = torch.zeros((6, 2), dtype=torch.int64, device="cuda")
    preds = ops.accuracy_counters(logits, answers, counts, tokens=question, token_to_cat=table.cuda(), want_preds=True)
    assert torch.equal(preds.long(), logits.argmax(1)) and int(preds[5]) == 3
    # the reference's Python loop (validate.py:59-80, 136-160 pattern)
    agree = (logits.argmax(1) == answers).cpu()
    ref = np.zeros((6, 2), dtype=np.int64)
    key = question[:, 0].cpu()
    idx_to_token = {v: k for k, v in words.items()}
    for i, w in enumerate(key):
        name = idx_to_token.get(int(w))
        if name is not None:
            c = ("what", "who", "how", "when", "where").index(name)
            ref[c, 0] += int(agree[i]); ref[c, 1] += 1
    ref[5] = (int(agree.sum()), B)
    assert np.array_equal(counts.cpu().numpy(), ref)
    # SVQA: explicit category ids, accumulating over two calls
    cat = torch.randint(0, 15, (B,), generator=g).cuda()
    counts = torch.zeros((16, 2), dtype=torch.int64, device="cuda")
    ops.accuracy_counters(logits, answers, counts, category=cat)
    ops.accuracy_counters(logits, answers, counts, category=cat)
    for c in range(15):
        sel = (cat == c).cpu()
        assert int(counts[c, 1]) == 2 * int(sel.sum()) and int(counts[c, 0]) == 2 * int(agree[sel].sum())
    assert int(counts[15, 1]) == 2 * B


def test_eval_engine_fast_path_and_graph_replay():
    from dualvgr_videoqa_b200.evaluate import EvalEngine, SVQA_CATEGORIES
    cfg = (12, 20, 8, 32, 60, 2)
    model = _build(cfg)
    app, mot, q, qlen, ans = [t.cuda() for t in orc.make_inputs(*cfg[:5])]
    cat = (torch.arange(cfg[0]) % 15).cuda()
    ev = EvalEngine(model, SVQA_CATEGORIES)
    logits, preds = ev.step(app, mot, q, qlen, ans, cat, want_preds=True)
    # eval mode: deterministic, running BatchNorm statistics, bf16 views instead of fp32 graph outputs
    with torch.no_grad():
        out = model.eval()(app, mot, q, qlen)
    assert torch.equal(out[0], logits) and out[3][0].dtype == torch.bfloat16
    r1 = ev.result()
    assert r1["counts"]["all"][1] == cfg[0] and r1["counts"]["all"][0] == int((logits.argmax(1) == ans).sum())
    # the captured eval step counts exactly like the eager one
    ev.reset()
    ev.capture(app, mot, q, qlen, ans, cat)
    assert ev.result()["counts"]["all"][1] == 0
    ev.replay()
    ev.load_batch(app=app * 0.5)
    l2 = ev.replay().clone()
    r2 = ev.result()
    assert r2["counts"]["all"][1] == 2 * cfg[0] and not torch.equal(l2, logits)
    sd64 = orc.cast_state_dict(orc.make_state_dict(cfg[5], cfg[3], cfg[4]), torch.float64)
    ref = orc.dualvgr_forward(sd64, cfg[5], app.double().cpu(), mot.double().cpu(), q.cpu(), qlen.cpu(), training=False)[0]
    assert float((logits.double().cpu() - ref).norm() / ref.norm()) < 2e-2


def test_pinned_loader_feeds_engine_from_fp32_and_bf16_stores(tmp_path):
    from dualvgr_videoqa_b200.loader import FeatureStore, PinnedBatchLoader
    from dualvgr_videoqa_b200.engine import TrainEngine
    g = np.random.default_rng(0)
    Vd, N, F, Dv, S, L, A, Vq = 10, 8, 16, 2048, 23, 6, 10, 30
    app = np.abs(g.standard_normal((Vd, N, F, Dv), dtype=np.float32))
    mot = np.abs(g.standard_normal((Vd, N, Dv), dtype=np.float32))
    np.save(tmp_path / "app.npy", app); np.save(tmp_path / "mot.npy", mot)
    np.save(tmp_path / "app16.npy", FeatureStore.to_bf16_bits(app)); np.save(tmp_path / "mot16.npy", FeatureStore.to_bf16_bits(mot))
    assert np.array_equal(FeatureStore.to_bf16_bits(app[:2]).view(np.int16),
                          torch.from_numpy(app[:2]).to(torch.bfloat16).view(torch.int16).numpy())
    qlen = g.integers(2, L + 1, S)
    samples = {"video_idx": g.integers(0, Vd, S), "question": g.integers(2, Vq, (S, L)) * (np.arange(L)[None] < qlen[:, None]),
               "question_len": qlen, "answer": g.integers(0, A, S)}
    for name, dt in (("app.npy", torch.float32), ("app16.npy", torch.bfloat16)):
        store = FeatureStore.open(str(tmp_path / name), str(tmp_path / name.replace("app", "mot")))
        loader = PinnedBatchLoader(store, samples, batch_size=5, device="cuda", drop_last=False)
        seen = 0
        for b_app, b_mot, b_q, b_len, b_ans in loader:
            n = b_app.shape[0]
            idx = np.arange(seen, seen + n)
            ref_app = torch.from_numpy(app[samples["video_idx"][idx]]).to(dt)
            assert b_app.dtype == dt and torch.equal(b_app.cpu(), ref_app)
            assert torch.equal(b_mot.cpu(), torch.from_numpy(mot[samples["video_idx"][idx]]).to(dt))
            assert np.array_equal(b_q.cpu().numpy(), samples["question"][idx]) and np.array_equal(b_ans.cpu().numpy(), samples["answer"][idx])
            seen += n
        assert seen == S and len(loader) == 5
    # and it drives the train engine (bf16-stored features, the captured step's static buffers)
    model = _build((5, N, L, A, Vq, 1))
    eng = TrainEngine(model, lr=1e-4)
    loader = PinnedBatchLoader(store, samples, batch_size=5, device="cuda", drop_last=True, shuffle=True)
    losses = []
    for epoch in range(2):
        for batch in loader:
            if eng.graph is None:
                eng.capture(*batch, warmup=1)
            eng.load_batch(*batch)
            losses.append(float(eng.replay()))
    assert len(losses) == 8 and all(np.isfinite(losses))
    eng.close()


def test_reference_train_and_validate_loop_bodies_drive_the_mirror():
    """train.py:131-160 and validate.py:44-63 against the mirror, statement by statement."""
    import os
    import sys
    import dualvgr_videoqa_b200
    pkg_dir = os.path.dirname(os.path.abspath(dualvgr_videoqa_b200.__path__[0] + "/__init__.py"))
    if pkg_dir not in sys.path:
        sys.path.insert(0, pkg_dir)                      # what a reference checkout does: the mirror's directory ahead on sys.path
    import model.models as modelset                      # train.py:20 — resolves to the mirror package
    from utils import todevice, common_loss, loss_dependence      # train.py:17 `from utils import *`
    from torch import nn, optim
    cfgB = (8, 8, 6, 10, 30, 2)
    B, N, L, A, V, U = cfgB
    cfg = types.SimpleNamespace(alpha=1.0, beta=1e-8, model_type="DualVGR", gpu_id=0,
                                dataset=types.SimpleNamespace(name="svqa"),
                                train=types.SimpleNamespace(num_of_nodes=N, lr=1e-4, batch_size=B))
    device = "cuda"
    model = modelset.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1,
                             unit_layers=U).to(device)                                           # train.py:69
    model.load_state_dict(orc.make_state_dict(U, A, V), strict=True)
    optimizer = optim.Adam(model.parameters(), cfg.train.lr)                                      # train.py:85
    criterion = nn.CrossEntropyLoss().to(device)                                                  # train.py:121
    app, mot, q, qlen, ans = orc.make_inputs(B, N, L, A, V)
    cats = torch.arange(B) % 15
    batch = [torch.arange(B), torch.arange(B), cats, ans.unsqueeze(1), app, mot, q, qlen]          # DataLoader.py:84 (svqa)

    def batch_accuracy(predicted, true):                                                          # train.py:352-356
        predicted = predicted.detach().argmax(1)
        return predicted == true

    model.train()
    losses = []
    for i in range(3):
        _, _, question_categories, answers, *batch_input = [todevice(x, device) for x in batch]    # train.py:133-134
        answers = answers.cuda().squeeze()
        optimizer.zero_grad()
        logits, aq_embed, mq_embed, com_app, com_motion, aq_fusion, mq_fusion = model(*batch_input)
        loss = criterion(logits, answers)
        loss_dep = 0
        loss_com = 0
        temp = len(aq_fusion)
        for k in range(temp):
            loss_dep += (loss_dependence(aq_fusion[k].cuda(), com_app[k].cuda(), cfg.train.num_of_nodes)
                         + loss_dependence(mq_fusion[k].cuda(), com_motion[k].cuda(), cfg.train.num_of_nodes))
            loss_com += common_loss(com_app[k].cuda(), com_motion[k].cuda())
        loss = loss + cfg.alpha * loss_com / temp + cfg.beta * loss_dep / temp
        loss.backward()
        nn.utils.clip_grad_norm_(model.parameters(), max_norm=12)
        optimizer.step()
        aggreeings = batch_accuracy(logits, answers)
        losses.append(loss.item())
        assert aggreeings.shape == (B,) and all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    assert all(np.isfinite(losses)) and temp == U and tuple(logits.shape) == (B, A)
    # validate.py:24,44-61
    from dualvgr_videoqa_b200.evaluate import EvalEngine, SVQA_CATEGORIES
    model.eval()
    ev = EvalEngine(model, SVQA_CATEGORIES)
    with torch.no_grad():
        video_ids, question_ids, question_categories, answers, *batch_input = [todevice(x, device) for x in batch]
        answers = answers.to(device).squeeze()
        logits, aq_embed, mq_embed, com_app, com_motion, aq_fusion, mq_fusion = model(*batch_input)
        preds = logits.detach().argmax(1)
        agreeings = (preds == answers)
        ev.step(*batch_input, answers, question_categories)
    res = ev.result()
    assert res["counts"]["all"] == (int(agreeings.sum()), B)
    for c, name in enumerate(SVQA_CATEGORIES):                                                     # validate.py:97-130, on the device
        sel = question_categories == c
        assert res["counts"][name] == (int(agreeings[sel].sum()), int(sel.sum()))
    sd = model.state_dict()                                                                        # train.py:359-367 / validate.py:286
    model2 = modelset.DualVGR(vocab=orc.make_vocab(V, A), num_of_nodes=N, graph_module="GAT", graph_layers=1, unit_layers=U)
    model2.load_state_dict(sd, strict=True)
