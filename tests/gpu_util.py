"""Helpers shared by the -m gpu parity tests (they all call through the C ABI via fastvision_b200)."""
import numpy as np
import torch

from conftest import T


def cuda(x):
    if isinstance(x, np.ndarray):
        x = T(x)
    return x.cuda()


def close(a, b, rtol=1e-5, atol=1e-6):
    """rtol 1e-5 is the north-star tolerance for fp32 boxes / IoUs / losses (BASELINE.json)."""
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def nms_matches_or_excused(got_rows, want_rows, boxes, scores, thr):
    """Keep lists must be identical unless some evaluated pair has |IoU - thr| < 1e-6 (north-star slack)."""
    if list(got_rows) == list(want_rows):
        return True
    from oracle.nms import nms_greedy
    _, margin = nms_greedy(boxes, scores, thr, return_iou_margin=True)
    return margin < 1e-6
