"""QuadrupletEvaluator paired distances against the sklearn-based oracle."""
import csv
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _quads(B, D, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, D, generator=g)
    p = a + 0.4 * torch.randn(B, D, generator=g)
    pp = a + 0.8 * torch.randn(B, D, generator=g)
    n = torch.randn(B, D, generator=g)
    return a, p, pp, n


@pytest.mark.parametrize("B,D", [(257, 384), (64, 768), (5, 7), (1, 1)])
def test_paired_distances_and_counts(B, D):
    import qst_b200
    from oracle import quad_eval_oracle as qo
    a, p, pp, n = _quads(B, D, 14 + B)
    counts, dist = qst_b200.paired_distance_counts(*[x.to(_dev()) for x in (a, p, pp, n)], want_distances=True)
    dist = dist.cpu().numpy()
    want = []
    for other in (p, pp, n):
        want.append(qo.paired_distances(a.numpy(), other.numpy()))
    for m in range(3):                       # cos, manhattan, euclid
        for k in range(3):                   # pos, part, neg
            np.testing.assert_allclose(dist[:, m * 3 + k], want[k][m], rtol=2e-5, atol=2e-6)
    # counts: identical to comparing the oracle's distances, except rows where the two distances
    # are within float rounding of each other
    for m in range(3):
        for j, (x, y) in enumerate(((0, 1), (0, 2), (1, 2))):
            dx, dy = want[x][m], want[y][m]
            sure = np.abs(dx - dy) > 1e-5 * np.maximum(np.abs(dx), np.abs(dy))
            lo = int(((dx < dy) & sure).sum())
            hi = lo + int((~sure).sum())
            assert lo <= int(counts[m, j]) <= hi, (m, j, lo, int(counts[m, j]), hi)


def test_evaluator_protocol_and_csv(tmp_path):
    import qst_b200
    from oracle import quad_eval_oracle as qo
    B, D = 300, 96
    a, p, pp, n = _quads(B, D, 3)
    table = torch.cat([a, p, pp, n]).to(_dev())
    model = qst_b200.synth.TableModel(table)
    ids = [[str(k * B + i) for i in range(B)] for k in range(4)]
    for main, key in ((None, None), (qst_b200.SimilarityFunction.COSINE, "cos"),
                      (qst_b200.SimilarityFunction.MANHATTAN, "manhattan"),
                      (qst_b200.SimilarityFunction.EUCLIDEAN, "euclid")):
        ev = qst_b200.QuadrupletEvaluator(*ids, gamma=0.6, main_distance_function=main, name="t")
        got = ev(model, output_path=str(tmp_path), epoch=2, steps=7)
        want, parts = qo.quadruplet_accuracy(a.numpy(), p.numpy(), pp.numpy(), n.numpy(), gamma=0.6, main=key)
        assert abs(got - want) <= 2.0 / B, (main, got, want)
        for k, v in parts.items():
            assert abs(ev.last_accuracies[k] - v) <= 1.0 / B
    rows = list(csv.reader(open(os.path.join(tmp_path, "quadruplet_evaluation_t_results.csv"))))
    assert rows[0] == ["epoch", "steps", "pos_part_accuracy", "pos_neg_accuracy", "part_neg_accuracy", "global_accuracy"]
    assert len(rows) == 5 and rows[1][:2] == ["2", "7"]
    trip = list(csv.reader(open(os.path.join(tmp_path, "triplet_evaluation_pos_neg_results.csv"))))
    assert trip[0] == ["epoch", "steps", "accuracy_cosinus", "accuracy_manhattan", "accuracy_euclidean"]
    # sampling from dataset-style items (dict with lists) and InputExample-like objects
    class Ex:
        def __init__(self, t):
            self.texts = t
    items = [{"reference": "0", "positive": ["300", "301"], "part_positive": "600", "negative": ["900"]},
             (Ex(["1", "301", "601", "901"]), 0)]
    ev2 = qst_b200.QuadrupletEvaluator.from_input_examples(items, gamma=0.5)
    assert ev2.anchors == ["0", "1"] and ev2.negatives == ["900", "901"] and ev2.positives[0] in ("300", "301")
    assert 0.0 <= ev2(model) <= 1.0


@pytest.mark.parametrize("n,batch,use_amp", [(150, 32, False), (64, 64, False), (33, 8, True), (5, 32, False)])
def test_loss_evaluator_running_mean_and_log(tmp_path, n, batch, use_amp):
    """QuadrupletLossEvaluator (models/evaluators.py:34-128): per-batch fused losses vs the loss oracle
    (1e-5 relative), running mean bit-identical to the reference expression on the same batch values,
    JSON log appended per call."""
    import json
    import qst_b200
    from oracle import loss_eval_oracle
    D = 96
    a, p, pp, neg = _quads(n, D, 40 + n)
    model = qst_b200.synth.TableModel(torch.cat([a, p, pp, neg]).to(_dev()))

    class Example:
        def __init__(self, texts):
            self.texts = texts

    dataset = [Example([str(k * n + i) for k in range(4)]) for i in range(n)]
    loss = qst_b200.GammaQuadrupletLoss(gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5)
    ev = qst_b200.QuadrupletLossEvaluator(dataset, loss, batch_size=batch, use_amp=use_amp)
    out = str(tmp_path)
    got = ev(model, output_path=out, epoch=1, steps=10)
    want_avg, want_losses = loss_eval_oracle.evaluate(a, p, pp, neg, batch, gamma=0.6, margin_pos_neg=1.0,
                                                      margin_pos_part=0.5, margin_part_neg=0.5)
    np.testing.assert_allclose(ev.last_batch_losses, want_losses.numpy(), rtol=1e-5)
    assert got == float(loss_eval_oracle.running_average(list(torch.from_numpy(ev.last_batch_losses))))
    assert abs(got - float(want_avg)) <= 1e-5 * abs(float(want_avg))
    again = ev(model, output_path=out, epoch=2, steps=-1)
    assert again == got
    with open(os.path.join(out, "_quadruplet_loss_eval.json")) as fp:
        log = json.load(fp)
    assert log == {"epoch": [1, 2], "steps": [10, -1], "average_loss": [got, got]}
    assert ev(model) == got                      # no output path: nothing written, same value


def test_dissimilar_mask_matches_cos_sim_threshold():
    """Negative-mining filter (dataset/quadruplet_dataset.py:229-234) on the device scorer."""
    import qst_b200
    from oracle import ir_oracle
    g = torch.Generator().manual_seed(5)
    ref = torch.randn(384, generator=g)
    cand = torch.cat([ref[None] * 0.5 + 0.1 * torch.randn(10, 384, generator=g), torch.randn(40, 384, generator=g)])
    mask, scores = qst_b200.dissimilar_mask(ref.to(_dev()), cand.to(_dev()), 0.2)
    want = ir_oracle.cos_sim(ref, cand)[0]
    np.testing.assert_allclose(scores.cpu().numpy(), want.numpy(), atol=2e-6)
    sure = (want - 0.2).abs() > 1e-5
    assert torch.equal(mask.cpu()[sure], (want <= 0.2)[sure])
    assert int(mask.sum()) >= 35 and not bool(mask[:10].any())


def test_device_paths_against_the_reference_codes_recorded_outputs(tmp_path):
    """tests/golden/evaluators_golden.json holds what the reference's OWN code produced
    (tests/golden/make_evaluators_golden.py): ``euclidean_score`` of models/evaluators.py:392-405 on a seeded
    case, and the incremental mean of ``QuadrupletLossEvaluator`` (:84-98) over scripted batch losses.  Here
    the device paths: the drop-in's ``euclidean_score`` (K3 arithmetic) against the recorded matrix, and the
    drop-in ``QuadrupletLossEvaluator`` run end to end on the GPU with a loss module that returns the scripted
    values -- same running mean, same JSON log."""
    import json
    import qst_b200
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "evaluators_golden.json")) as fp:
        golden = json.load(fp)
    g = golden["euclidean_score"]
    a, b = torch.tensor(g["a"]), torch.tensor(g["b"])
    got = qst_b200.euclidean_score(a.to(_dev()), b.to(_dev())).cpu()
    torch.testing.assert_close(got, torch.tensor(g["scores"]), rtol=1e-5, atol=1e-6)     # cdist vs direct (q-c)^2 sums
    one = qst_b200.euclidean_score(a[0].to(_dev()), b[1].to(_dev())).cpu()
    torch.testing.assert_close(one, torch.tensor(g["one_d"]), rtol=1e-5, atol=1e-6)

    class ScriptedLoss(torch.nn.Module):
        """Returns the scripted batch losses as device scalars (the fused loss itself has its own tests)."""

        def __init__(self, values):
            super().__init__()
            self.values = list(values)

        def forward(self, x_anchor, x_pos, x_part, x_neg, **kw):
            assert x_anchor.is_cuda and x_anchor.shape == x_neg.shape
            return torch.tensor(self.values.pop(0), dtype=torch.float32, device=x_anchor.device)

    table = torch.randn(64, 16, generator=torch.Generator().manual_seed(5)).to(_dev())
    model = qst_b200.synth.TableModel(table)
    for case in golden["loss_evaluator"]:
        n_batches = len(case["batch_sizes_seen"])
        losses = golden["batch_losses"][:n_batches]
        items = [(str(i % 64), str((i + 1) % 64), str((i + 2) % 64), str((i + 3) % 64)) for i in range(case["n_items"])]
        out = tmp_path / f"{case['n_items']}_{case['batch_size']}"
        out.mkdir()
        loss = ScriptedLoss(losses + losses[::-1])
        ev = qst_b200.QuadrupletLossEvaluator(items, loss, batch_size=case["batch_size"])
        first = ev(model, output_path=str(out), epoch=0, steps=-1)
        second = ev(model, output_path=str(out), epoch=1, steps=40)
        assert [first, second] == case["returned"]
        assert open(out / "_quadruplet_loss_eval.json").read() == case["json_text"]
