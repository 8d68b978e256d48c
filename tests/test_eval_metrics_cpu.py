"""crab_b200/eval/metrics.py against the REFERENCE's own metric functions (fixtures generated from /root/reference/utils by
oracle/make_metrics_golden.py): AVVP segment / event level F-scores, AVS mask IoU / F-measure / null metric, AVSS per-class
mIoU / F-score, box IoU / cIoU."""
from pathlib import Path

import numpy as np
import pytest
import torch

from crab_b200.eval import metrics as M

G = torch.load(Path(__file__).resolve().parent / "golden" / "metrics.pt", weights_only=False)


@pytest.mark.parametrize("case", range(len(G["avvp"])))
def test_avvp_f_scores(case):
    c = G["avvp"][case]
    assert np.allclose(M.avvp_segment_level(*c["inputs"]), c["segment"], rtol=0, atol=1e-12)
    assert np.allclose(M.avvp_event_level(*c["inputs"]), c["event"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("case", range(len(G["mask"])))
def test_mask_metrics(case):
    c = G["mask"][case]
    assert torch.allclose(M.mask_iou(c["pred"], c["gt"]), c["iou"], rtol=1e-6, atol=0)
    assert abs(M.f_measure(c["pred"], c["gt"]) - c["f"]) < 1e-6
    assert torch.allclose(M.null_metric_s(c["pred"][:1]), c["s"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("case", range(len(G["avss"])))
def test_avss_miou_fscore(case):
    c = G["avss"][case]
    miou, fs, cnt, per = M.avss_miou_fscore(c["pred"], c["target"], T=2)
    assert torch.allclose(miou, c["miou"], rtol=1e-6, atol=1e-7) and torch.allclose(fs, c["fscore"], rtol=1e-6, atol=1e-7)
    assert torch.equal(cnt, c["cls_count"]) and torch.allclose(torch.stack(per), c["per_frame"], rtol=1e-6, atol=1e-7)


def test_box_iou_ciou():
    for c in G["box"]:
        assert abs(M.box_iou(c["a"], c["b"]) - c["iou"]) < 1e-9
        assert abs(M.box_ciou(c["a"], c["b"]) - c["ciou"]) < 1e-9


def test_event_extraction_edge_cases():
    assert M._runs(np.array([1, 1, 0, 1, 0, 0, 0, 1, 1, 1])) == [(0, 2), (3, 4), (7, 10)]
    assert M._runs(np.zeros(10)) == [] and M._runs(np.ones(10)) == [(0, 10)]
