"""evaluation metrics.  The session metrics are pinned against the reference's own
metrics/metrics.py (tests/golden/metrics.json); the streaming tf.metrics restatement against its definition."""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_offline_metrics_match_reference_implementation():
    from cikm2020_dmt_b200 import metrics as M
    with open(os.path.join(GOLD, "metrics.json")) as fh:
        g = json.load(fh)
    headers = [h.encode() for h in g["headers"]]
    sets, at = M.offline_metrics(g["schema"], headers, g["scores"])
    assert at == g["at_list"]
    for a in (M.CLICK, M.ORDER):
        assert np.allclose(sets[a][0], g["pre"][str(a)], rtol=0, atol=1e-12)
        assert np.allclose(sets[a][1], g["mrr"][str(a)], rtol=0, atol=1e-12)
    auc = M.offline_metrics_auc(g["schema"], headers, g["scores"])
    for a in (M.CLICK, M.ORDER):
        assert abs(float(auc[a][0]) - g["auc"][str(a)]) < 1e-12


def _tf_auc_bruteforce(labels, scores, k=200):
    eps = 1e-7
    th = [0.0 - eps] + [(i + 1) / (k - 1) for i in range(k - 2)] + [1.0 + eps]
    tpr, fpr = [], []
    for t in th:
        pred = scores > t
        tp, fp = np.sum(pred & labels), np.sum(pred & ~labels)
        fn, tn = np.sum(~pred & labels), np.sum(~pred & ~labels)
        tpr.append((tp + eps) / (tp + fn + eps))
        fpr.append(fp / (fp + tn + eps))
    tpr, fpr = np.asarray(tpr), np.asarray(fpr)
    return float(np.sum((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2))


def test_streaming_metrics_equal_definition_and_merge_over_batches():
    from cikm2020_dmt_b200 import metrics as M
    rng = np.random.default_rng(5)
    labels = rng.random(5000) < 0.07
    scores = np.clip(rng.normal(0.3, 0.2, 5000) + 0.25 * labels, 0, 1)
    scores[:10] = [0.0, 1.0, 0.5, 1 / 199, 2 / 199, 198 / 199, 0.5000001, 0.4999999, 0.25, 0.75]   # threshold edges
    m = M.StreamingBinaryMetrics()
    for lo in range(0, 5000, 700):                                     # streaming: counts accumulate over batches
        m.update(torch.from_numpy(labels[lo:lo + 700]).float(), torch.from_numpy(scores[lo:lo + 700]))
    r = m.result()
    assert abs(r["auc"] - _tf_auc_bruteforce(labels, scores)) < 1e-12
    pred = scores > 0.5
    assert abs(r["precision"] - np.sum(pred & labels) / np.sum(pred)) < 1e-12
    assert abs(r["recall"] - np.sum(pred & labels) / np.sum(labels)) < 1e-12
    # 200 thresholds discretise the exact ROC AUC to ~1e-3
    order = np.argsort(scores)
    ranks = np.empty(5000)
    ranks[order] = np.arange(1, 5001)
    exact = (ranks[labels].sum() - labels.sum() * (labels.sum() + 1) / 2) / (labels.sum() * (~labels).sum())
    assert abs(r["auc"] - exact) < 5e-3


def test_click_order_labels_from_mask():
    from cikm2020_dmt_b200 import metrics as M
    mask = torch.eye(5)
    clk, od = M.click_order_labels(mask)
    assert clk.tolist() == [0, 1, 1, 1, 1] and od.tolist() == [0, 0, 0, 1, 1]     # labels 0,1,2,4,5 (run_dnn.py:221,231)
    empty = M.StreamingBinaryMetrics().result()
    assert empty["precision"] == 0.0 and empty["recall"] == 0.0
