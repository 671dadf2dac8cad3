"""CPU test of the host half of the device-side evaluate: metrics from per-class counts must equal
sklearn's accuracy_score / f1_score on the same indicator matrices (reference model/evaluation.py:8-12)."""
import numpy as np
from sklearn.metrics import accuracy_score, f1_score

from rgcn_b200.evaluation import metrics_from_counts


def _counts(pred, y):
    tp = ((pred == 1) & (y == 1)).sum(0)
    fp = ((pred == 1) & (y == 0)).sum(0)
    fn = ((pred == 0) & (y == 1)).sum(0)
    exact = (pred == y).all(1).sum()
    return np.concatenate([tp, fp, fn, [exact]]).astype(np.int64)


def test_metrics_from_counts_equal_sklearn():
    rng = np.random.default_rng(0)
    for n, c, p in ((50, 4, 0.3), (200, 26, 0.1), (30, 2, 0.5), (10, 3, 0.0)):
        y = (rng.random((n, c)) < max(p, 0.05)).astype(np.int64)
        pred = (rng.random((n, c)) < p).astype(np.int64)      # p = 0: nothing predicted -> zero_division path
        got = metrics_from_counts(_counts(pred, y), n)
        want = (accuracy_score(y, pred), f1_score(y, pred, average='weighted', zero_division=0),
                f1_score(y, pred, average='macro', zero_division=0))
        assert np.allclose(got, want, atol=1e-12), (got, want)
