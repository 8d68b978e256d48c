"""TEST INFRASTRUCTURE — golden vectors for crab_b200/eval/metrics.py: run the REFERENCE's own metric functions
(/root/reference/utils/{avvp_eval_metrics,avss_utils,ciou}.py) on seeded random inputs and record inputs + outputs.
    python -m oracle.make_metrics_golden      # writes tests/golden/metrics.pt (~0.3 MB)"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/utils")


def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, REF / f"{name}.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    if not REF.exists():
        raise SystemExit("reference checkout not found")
    avvp, avss, ciou = load("avvp_eval_metrics"), load("avss_utils"), load("ciou")
    rng = np.random.default_rng(5)
    out = {"avvp": [], "mask": [], "avss": [], "box": []}
    for case in range(12):
        dens = [0.05, 0.15, 0.3, 0.5][case % 4]
        arrs = [(rng.random((25, 10)) < dens).astype(np.float64) for _ in range(6)]
        if case == 3:
            arrs = [np.zeros((25, 10)) for _ in range(6)]          # all true negatives
        if case == 7:
            arrs[3:] = [a.copy() for a in arrs[:3]]                # perfect prediction
        out["avvp"].append({"inputs": arrs, "segment": avvp.segment_level(*arrs), "event": avvp.event_level(*arrs)})
    g = torch.Generator().manual_seed(9)
    for case in range(6):
        n, h, w = 3, 56, 56
        gt = (torch.rand(n, h, w, generator=g) < 0.3).float()
        if case % 2:
            gt[1] = 0
        pred = 4 * (gt - 0.5) + 2.5 * torch.randn(n, h, w, generator=g)
        out["mask"].append({"pred": pred, "gt": gt, "iou": avss.mask_iou(pred, gt), "f": avss.Eval_Fmeasure(pred, gt),
                            "s": avss.metric_s_for_null(pred[:1])})
    for case in range(4):
        bf, c, h, w = 4, 7, 24, 24
        tgt = torch.randint(0, c, (bf, h, w), generator=g)
        pred = 3 * torch.nn.functional.one_hot(tgt, c).permute(0, 3, 1, 2).float() + 2 * torch.randn(bf, c, h, w, generator=g)
        miou, fs, cnt, per = avss.calc_color_miou_fscore(pred, tgt, T=2)
        out["avss"].append({"pred": pred, "target": tgt, "miou": miou, "fscore": fs, "cls_count": cnt, "per_frame": torch.stack(per)})
    for case in range(16):
        b = []
        for _ in range(2):
            x0, y0 = rng.uniform(0, 60, 2)
            b.append([float(x0), float(y0), float(x0 + rng.uniform(5, 60)), float(y0 + rng.uniform(5, 60))])
        out["box"].append({"a": b[0], "b": b[1], "iou": float(ciou.intersection_over_union(b[0], b[1])), "ciou": float(ciou.c_iou(b[0], b[1]))})
    p = ROOT / "tests" / "golden" / "metrics.pt"
    torch.save(out, p)
    print(f"wrote {p} ({p.stat().st_size / 1e3:.0f} KB)")


if __name__ == "__main__":
    main()
