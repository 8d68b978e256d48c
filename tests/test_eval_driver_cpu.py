"""The batched-evaluation loops (crab_b200/eval/driver.py) on a stand-in model / tokenizer: call pattern of
scripts/finetune/inference_hyper_lora.py (generate(**sample, use_cache=True, max_new_tokens=500) on the loader's batches,
batch_decode, one record per sample; generate_avs + running mask IoU / F-measure; avss class means)."""
import json

import torch

from crab_b200.eval import driver as D
from crab_b200.eval import metrics as M


class Tok:
    def batch_decode(self, ids, skip_special_tokens=False):
        return [" ".join(str(int(t)) for t in row if not (skip_special_tokens and int(t) >= 100)) for row in ids]

    def decode(self, ids, skip_special_tokens=False):
        return self.batch_decode([ids], skip_special_tokens)[0]


class Model:
    def __init__(self):
        self.calls = []

    def generate(self, batch_input_ids, batch_labels, batch_X_modals, batch_task_names, use_cache, max_new_tokens):
        self.calls.append((len(batch_input_ids), use_cache, max_new_tokens, batch_input_ids[0].device.type))
        return torch.stack([ids[:3] + 100 for ids in batch_input_ids])

    def generate_avs(self, batch_input_ids, batch_labels, batch_X_modals, batch_task_names, use_cache, max_new_tokens):
        self.calls.append(("avs", max_new_tokens))
        gt = batch_X_modals[0]["<mask>"]
        if gt.dtype == torch.long:                                            # avss: class scores
            pred = 4 * torch.nn.functional.one_hot(gt[0], 5).permute(2, 0, 1).float()
            pred[:, :4, :4] = 0
            pred[0, :4, :4] = 4
        else:
            pred = 8 * (gt - 0.5)
            pred[:, :8, :8] = -4                                               # some wrong pixels
        return {"output_ids": batch_input_ids[0][None, :2], "pred_masks": [pred]}


def loader(n_batches, bs):
    for b in range(n_batches):
        ids = [torch.arange(5) + 10 * (b * bs + i) for i in range(bs)]
        yield {"batch_input_ids": ids, "batch_labels": [i.clone() for i in ids], "batch_X_modals": [{} for _ in ids],
               "batch_task_names": ["avqa"] * bs, "batch_metadata": [{"vid": b * bs + i, "output": "yes"} for i in range(bs)]}


def test_text_task_call_pattern(tmp_path):
    m, fp = Model(), tmp_path / "infer_results.jsonl"
    recs = D.run_text_task(m, Tok(), loader(3, 8), task="avqa", out_path=str(fp), device="cpu")
    assert m.calls == [(8, True, 500, "cpu")] * 3 and len(recs) == 24
    lines = [json.loads(x) for x in fp.read_text().splitlines()]
    assert lines[9] == {"vid": 9, "output": "yes", "predict": "190 191 192"}
    recs = D.run_text_task(m, Tok(), loader(1, 2), task="avvp", device="cpu")     # avvp decodes with skip_special_tokens=True
    assert recs[0]["predict"] == ""


def test_avs_and_avss_tasks():
    g = torch.Generator().manual_seed(0)
    gt = (torch.rand(1, 32, 32, generator=g) < 0.4).float()
    sample = {"batch_input_ids": [torch.arange(4)], "batch_labels": [torch.arange(4)], "batch_X_modals": [{"<mask>": gt}],
              "batch_task_names": ["s4"], "batch_metadata": [{"instruction": "q", "output": "a"}]}
    m = Model()
    r = D.run_avs_task(m, Tok(), [sample, sample], task="s4", device="cpu")
    pred = 8 * (gt - 0.5)
    pred[:, :8, :8] = -4
    assert r["count"] == 2 and abs(r["miou"] - float(M.mask_iou(pred, gt))) < 1e-6 and abs(r["fscore"] - M.f_measure(pred, gt)) < 1e-6
    assert m.calls == [("avs", 100)] * 2
    r0 = D.run_avs_task(Model(), Tok(), [sample], task="ref_avs", device="cpu", null_split=True)
    assert abs(r0["s"] - float(M.null_metric_s(pred))) < 1e-6
    lab = torch.randint(0, 5, (1, 16, 16), generator=g)
    s2 = dict(sample, batch_X_modals=[{"<mask>": lab}])
    r2 = D.run_avss_task(Model(), Tok(), [s2], n_classes=5, device="cpu")
    assert 0.5 < r2["miou"] <= 1.0 and 0.5 < r2["f_score"] <= 1.0 and len(r2["records"]) == 1
