"""Jaccard, the reference's training metric (deeplabv3p/metrics.py:30-45): host arithmetic vs the literal restatement (CPU), and the
device pixel counts vs numpy (GPU, integer: bit exact)."""
import numpy as np
import pytest

from oracle import head_ref as R


def _case(seed, B=3, H=37, W=29, NC=5, ignore_frac=0.1, drop_class=None):
    rng = np.random.default_rng(seed)
    gt = rng.integers(0, NC, size=(B, H, W)).astype(np.uint8)
    pred = np.where(rng.random(gt.shape) < 0.7, gt, rng.integers(0, NC, size=gt.shape)).astype(np.uint8)
    gt[rng.random(gt.shape) < ignore_frac] = 255
    if drop_class is not None:
        gt[0][gt[0] == drop_class] = (drop_class + 1) % NC        # image 0 does not contain the class: not a "legal batch" for it
    return pred, gt


def _counts_numpy(pred, gt, NC):
    B = pred.shape[0]
    c = np.zeros((B, 3, NC + 1), np.int64)
    for b in range(B):
        for i in range(NC + 1):
            t, p = gt[b] == i, pred[b] == i
            c[b, 0, i], c[b, 1, i], c[b, 2, i] = (t & p).sum(), t.sum(), p.sum()
    return c


@pytest.mark.parametrize('seed,drop', [(1, None), (2, 3), (3, 0)])
def test_jaccard_from_counts_equals_the_literal_restatement(seed, drop):
    from dlv3p_b200 import metrics
    pred, gt = _case(seed, drop_class=drop)
    ref = R.jaccard_metric(gt, pred, 5)
    got = metrics.jaccard_from_counts(_counts_numpy(pred, gt, 5))
    assert got == pytest.approx(ref, rel=0, abs=1e-15)
    assert 0.3 < ref < 1.0
    # a class absent from every image drops out of the mean; perfect prediction on the non-ignored pixels is not 1.0 because the
    # ignored pixels still count for the class they were predicted as (metrics.py:36-38)
    perfect = np.where(gt == 255, 0, gt).astype(np.uint8)
    assert R.jaccard_metric(gt, perfect, 5) < 1.0
    assert R.jaccard_metric(perfect, perfect, 5) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize('B,H,W,NC', [(2, 512, 512, 21), (3, 37, 29, 5), (1, 7, 5, 2), (8, 128, 128, 19)])
def test_jaccard_counts_on_the_device_are_bit_exact(gpu, B, H, W, NC):
    from dlv3p_b200 import metrics
    pred, gt = _case(B * 31 + NC, B=B, H=H, W=W, NC=NC)
    counts = metrics.jaccard_counts(pred, gt, NC)
    assert np.array_equal(counts, _counts_numpy(pred, gt, NC))
    assert metrics.jaccard(pred, gt, NC) == pytest.approx(R.jaccard_metric(gt, pred, NC), rel=0, abs=1e-15)
