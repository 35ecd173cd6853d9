"""Host side of the device metrics (gan_aug_pfa_b200.metrics): the ratio formulas of evaluate.py:47-64 from integer
confusion counts, checked against a restatement of calculate_metrics on random masks (CPU only; the counting kernel
itself is covered by tests/test_gpu_siamese.py)."""
import pytest
import torch

from gan_aug_pfa_b200 import metrics
from oracle import pix2pix_oracle as O


_calculate_metrics = O.calculate_metrics


@pytest.mark.parametrize("seed,p_pos,p_pred", [(0, 0.05, 0.07), (1, 0.5, 0.5), (2, 0.0, 0.1), (3, 0.2, 0.0), (4, 1.0, 1.0)])
def test_metrics_from_counts_matches_calculate_metrics(seed, p_pos, p_pred):
    g = torch.Generator().manual_seed(seed)
    targets = (torch.rand(64, 64, generator=g) < p_pos).float()
    probs = torch.where(torch.rand(64, 64, generator=g) < p_pred, torch.tensor(0.9), torch.tensor(0.1))
    ref, (tp, fp, fn, tn) = _calculate_metrics(probs, targets)
    got = metrics.metrics_from_counts(tp, fp, fn, tn)
    assert set(got) == set(metrics.METRIC_KEYS)
    for k in metrics.METRIC_KEYS:
        assert got[k] == pytest.approx(ref[k], rel=1e-6, abs=1e-9)


def test_metrics_module_has_no_cpu_counting_path():
    """The counting itself only exists as a CUDA kernel: CPU tensors are rejected, not silently handled."""
    with pytest.raises((ValueError, RuntimeError)):
        metrics.confusion_counts(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64))
