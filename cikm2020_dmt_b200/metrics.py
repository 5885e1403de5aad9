"""Evaluation metrics in front of / behind the hot path, without TensorFlow / pandas.

* `StreamingBinaryMetrics` -- `tf.metrics.precision`, `tf.metrics.recall` and `tf.metrics.auc` as `run_dnn.py`
  uses them (:217-241 train, :700-720 test): click label = sum(mask[:, 1:5]), order label = mask[:, 3] + mask[:, 4];
  predictions thresholded at 0.5; AUC = TF-1's default 200-threshold trapezoidal ROC with streaming confusion
  counts (thresholds (i+1)/199 bracketed by -1e-7 and 1+1e-7, epsilon 1e-7 in the rates).  The counts are plain
  tensors: data-parallel ranks allreduce them (`merge`) exactly like the metric variables of a TF session.
* `offline_metrics` / `offline_metrics_auc` -- metrics/metrics.py:90-276: session-grouped Precision@N and MRR@N
  (N in 2..14) for click (label >= 2) and order (label >= 5) on score = p_ctr + p_cvr (run_dnn.py:847-849), and the
  per-uuid ROC AUC averaged over the users with at least two impressions (a one-class user counts as 1, as the
  reference's try/except does).  Pinned against the reference's own metrics.py (tests/golden/make_golden_metrics.py).
"""
from typing import Dict, List, Sequence

import numpy as np
import torch

CLICK, ORDER = 2, 5                      # metrics/metrics.py:46-47
AT_LIST = [2, 4, 6, 8, 10, 12, 14]       # metrics/metrics.py:49
_EPS = 1e-7


class StreamingBinaryMetrics(object):
    def __init__(self, num_thresholds=200, device="cpu"):
        k = num_thresholds
        th = [(i + 1) * 1.0 / (k - 1) for i in range(k - 2)]
        self.thresholds = torch.tensor([0.0 - _EPS] + th + [1.0 + _EPS], dtype=torch.float64, device=device)
        self.pos_hist = torch.zeros(k + 1, dtype=torch.float64, device=device)   # by number of thresholds < score
        self.neg_hist = torch.zeros(k + 1, dtype=torch.float64, device=device)
        self.conf = torch.zeros(4, dtype=torch.float64, device=device)           # tp, fp, fn, tn at 0.5

    def update(self, labels: torch.Tensor, scores: torch.Tensor):
        """labels in {0, 1} (any float / bool tensor), scores = sigmoid outputs, same shape."""
        labels = labels.reshape(-1).to(self.pos_hist.device) > 0.5
        scores = scores.reshape(-1).to(self.pos_hist.device, torch.float64)
        # tf.metrics.auc: prediction is positive at threshold t iff score > t
        bucket = torch.bucketize(scores, self.thresholds, right=False)            # thresholds strictly below score
        k1 = self.pos_hist.numel()
        self.pos_hist += torch.bincount(bucket[labels], minlength=k1).to(torch.float64)
        self.neg_hist += torch.bincount(bucket[~labels], minlength=k1).to(torch.float64)
        pred = scores > 0.5
        self.conf += torch.stack([(pred & labels).sum(), (pred & ~labels).sum(), (~pred & labels).sum(),
                                  (~pred & ~labels).sum()]).to(torch.float64)

    def merge(self, group=None):
        """Sum the counts over the data-parallel ranks."""
        import torch.distributed as dist
        for t in (self.pos_hist, self.neg_hist, self.conf):
            dist.all_reduce(t, group=group)

    def result(self) -> Dict[str, float]:
        pos, neg = self.pos_hist.cpu(), self.neg_hist.cpu()
        # tp[i] = positives with score > thresholds[i] = positives whose bucket (count of thresholds below) > i
        tp = torch.flip(torch.cumsum(torch.flip(pos, [0]), 0), [0])[1:]
        fp = torch.flip(torch.cumsum(torch.flip(neg, [0]), 0), [0])[1:]
        fn, tn = pos.sum() - tp, neg.sum() - fp
        tpr = (tp + _EPS) / (tp + fn + _EPS)
        fpr = fp / (fp + tn + _EPS)
        auc = float(((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2.0).sum())
        tp5, fp5, fn5, _ = [float(v) for v in self.conf.cpu()]
        return {"auc": auc,
                "precision": tp5 / (tp5 + fp5) if tp5 + fp5 > 0 else 0.0,      # tf.metrics.precision: 0 when empty
                "recall": tp5 / (tp5 + fn5) if tp5 + fn5 > 0 else 0.0}


def click_order_labels(mask: torch.Tensor):
    """run_dnn.py:221,231: click = sum(mask[:, 1:5]), order = mask[:, 3] + mask[:, 4]."""
    return mask[:, 1:5].sum(-1), mask[:, 3] + mask[:, 4]


# ----------------------------------------------------------------------------- metrics/metrics.py
def _columns(header_schema: Sequence[str], headers: Sequence[bytes]):
    rows = [(h.decode() if isinstance(h, (bytes, bytearray)) else h).strip().split("\t") for h in headers]
    idx = {name: i for i, name in enumerate(header_schema)}
    return rows, idx


def _groups(keys: List[str]):
    order: Dict[str, List[int]] = {}
    for i, k in enumerate(keys):
        order.setdefault(k, []).append(i)
    return [np.asarray(order[k]) for k in sorted(order)]      # pandas groupby sorts the keys


def offline_metrics(header_schema, headers, scores):
    """get_offline_metrics (metrics.py:111-199): {CLICK: (pre@N, mrr@N), ORDER: (pre@N, mrr@N)}, AT_LIST."""
    rows, idx = _columns(header_schema, headers)
    label = np.asarray([int(r[idx["label"]]) for r in rows])
    scores = np.asarray(scores, dtype=np.float64)
    groups = _groups([r[idx["sid"]] for r in rows])
    pre = {CLICK: np.zeros(len(AT_LIST)), ORDER: np.zeros(len(AT_LIST))}
    mrr = {CLICK: np.zeros(len(AT_LIST)), ORDER: np.zeros(len(AT_LIST))}
    for g in groups:
        # sort_values(by=['score', 'label'], ascending=[False, True])
        o = g[np.lexsort((label[g], -scores[g]))]
        for i, n in enumerate(AT_LIST):
            top = label[o[:n]]
            if top.size == 0:
                continue
            for action in (CLICK, ORDER):
                hit = top >= action
                pre[action][i] += hit.sum() * 1.0 / top.size
                first = np.flatnonzero(hit)
                if first.size:
                    mrr[action][i] += 1.0 / float(first[0] + 1)
    n_groups = max(len(groups), 1)
    return {a: (pre[a] / n_groups, mrr[a] / n_groups) for a in (CLICK, ORDER)}, list(AT_LIST)


def _roc_auc(y: np.ndarray, s: np.ndarray) -> float:
    """sklearn.metrics.roc_auc_score for binary y: Mann-Whitney U with average ranks for ties."""
    order = np.argsort(s, kind="mergesort")
    ranks = np.empty(len(s), dtype=np.float64)
    ss = s[order]
    i = 0
    while i < len(ss):
        j = i
        while j + 1 < len(ss) and ss[j + 1] == ss[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    n_pos = float(y.sum())
    n_neg = float(len(y) - n_pos)
    return (ranks[y > 0].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg)


def offline_metrics_auc(header_schema, headers, scores, group_method="uuid"):
    """get_offline_metrics_auc (metrics.py:204-276): {CLICK: mean per-group AUC, ORDER: ...} over the groups with
    more than one row; a group with a single class contributes 1 (the reference's `except: return 1`)."""
    rows, idx = _columns(header_schema, headers)
    label = np.asarray([int(r[idx["label"]]) for r in rows])
    scores = np.asarray(scores, dtype=np.float64)
    total = {CLICK: 0.0, ORDER: 0.0}
    valid = 0
    for g in _groups([r[idx[group_method]] for r in rows]):
        if len(g) == 1:
            continue
        valid += 1
        for action in (CLICK, ORDER):
            y = (label[g] >= action).astype(np.int64)
            total[action] += 1.0 if y.min() == y.max() else _roc_auc(y, scores[g])
    return {a: np.asarray([total[a] / valid if valid else float("nan")]) for a in (CLICK, ORDER)}
