"""Task loops of the reference's batched-evaluation script (scripts/finetune/inference_hyper_lora.py), restated as three small
generic loops over any dataloader that yields the reference's sample dicts
    {'batch_input_ids', 'batch_labels', 'batch_X_modals', 'batch_task_names', 'batch_metadata'}:

  run_text_task   — the next-token-prediction tasks (:158-479, :878-907: avqa, ave, avvp, next-qa, aok-vqa, arig, avcap and the
                    *_ntp variants): model.generate(**sample, use_cache=True, max_new_tokens=500) at whatever batch size the
                    loader yields (the reference uses 8), tokenizer.batch_decode, one jsonl record per sample with 'predict';
  run_avs_task    — s4 / ms3 / ref-avs (:593-823, :910-1080): model.generate_avs(max_new_tokens=100), sigmoid > 0.5 masks,
                    running mask IoU and F-measure (or the null metric S for the ref-avs "null" split);
  run_avss_task   — avss (:1138-1238): per-class IoU / F-score sums over frames, mean over classes with and without background.

Dataset construction, tokenisation and image dumps stay with the caller (they are file I/O outside the accelerated path); the
loops only need the five keys above, a tokenizer with batch_decode / decode, and a model with the reference's generate /
generate_avs signatures — crab_b200.models.unified_llama.UnifiedForCausalLM or the reference's own class.
"""
from __future__ import annotations

import json
from collections.abc import Mapping
from typing import Any, Callable, Dict, Iterable, List, Optional

import torch

from . import metrics as M

# task name -> (loop, max_new_tokens, skip_special_tokens) exactly as the reference calls them
TEXT_TASKS = {"avqa": (500, False), "ave": (500, False), "avvp": (500, True), "next_qa": (500, False), "aok_vqa": (500, False),
              "arig": (500, False), "avcap": (500, False), "s4_ntp": (500, False), "ms3_ntp": (100, False), "ref_avs_ntp": (100, False),
              "avss_ntp": (100, False)}
AVS_TASKS = {"s4": 100, "ms3": 100, "ref_avs": 100}


def prepare_sample(data: Any, device="cuda", dtype=None) -> Any:
    """Move every tensor of a nested sample to `device` (utils/util.py:33-47)."""
    if isinstance(data, Mapping):
        return type(data)({k: prepare_sample(v, device, dtype) for k, v in data.items()})
    if isinstance(data, (tuple, list)):
        return type(data)(prepare_sample(v, device, dtype) for v in data)
    if isinstance(data, torch.Tensor):
        return (data.to(dtype=dtype) if dtype is not None else data).to(device=device)
    return data


def write_jsonl(fp: Optional[str], record: Dict) -> None:
    if fp:
        with open(fp, "a") as f:
            f.write(json.dumps(record) + "\n")


@torch.no_grad()
def run_text_task(model, tokenizer, dataloader: Iterable[Dict], *, task: str = "avqa", out_path: Optional[str] = None,
                  device="cuda", max_new_tokens: Optional[int] = None, on_record: Optional[Callable[[Dict], None]] = None) -> List[Dict]:
    n_new, skip = TEXT_TASKS.get(task, (500, False))
    records = []
    for sample in dataloader:
        sample = dict(sample)
        meta = sample.pop("batch_metadata")
        sample = prepare_sample(sample, device)
        out = model.generate(**sample, use_cache=True, max_new_tokens=max_new_tokens or n_new)
        texts = tokenizer.batch_decode(out, skip_special_tokens=skip)
        for m, t in zip(meta, texts):
            rec = dict(m, predict=t)
            records.append(rec)
            write_jsonl(out_path, rec)
            if on_record:
                on_record(rec)
    return records


@torch.no_grad()
def run_avs_task(model, tokenizer, dataloader: Iterable[Dict], *, task: str = "s4", out_path: Optional[str] = None, device="cuda",
                 null_split: bool = False, max_new_tokens: Optional[int] = None) -> Dict[str, Any]:
    """Returns {'miou', 'fscore', 'count', 'records'} ('s' instead of the first two for the null split)."""
    n_new = max_new_tokens or AVS_TASKS.get(task, 100)
    tot_iou = tot_f = tot_s = 0.0
    count, records = 0, []
    for sample in dataloader:
        sample = dict(sample)
        meta = sample.pop("batch_metadata")
        gts = [x.get("<mask>") for x in sample["batch_X_modals"]]
        sample = prepare_sample(sample, device)
        res = model.generate_avs(**sample, use_cache=True, max_new_tokens=n_new)
        ids = res.get("output_ids")
        texts = tokenizer.batch_decode(ids, skip_special_tokens=False) if ids is not None else ["none"] * len(meta)
        masks = res.get("pred_masks")
        for i, m in enumerate(meta):
            rec = {"instruction": m.get("instruction"), "label": m.get("output"), "pred": texts[i] if i < len(texts) else "none"}
            if masks is not None and i < len(masks):
                pm = masks[i].detach().float().cpu()                       # (num_classes, H, W) logits
                if null_split:
                    s = float(M.null_metric_s(pm))
                    tot_s += s
                    rec["s"] = s
                else:
                    gt = gts[i].detach().float().cpu()
                    iou, f = float(M.mask_iou(pm, gt)), M.f_measure(pm, gt)
                    tot_iou += iou
                    tot_f += f
                    rec.update(iou=iou, fscore=f)
                count += 1
            records.append(rec)
            write_jsonl(out_path, rec)
    n = max(count, 1)
    if null_split:
        return {"s": tot_s / n, "count": count, "records": records}
    return {"miou": tot_iou / n, "fscore": tot_f / n, "count": count, "records": records}


@torch.no_grad()
def run_avss_task(model, tokenizer, dataloader: Iterable[Dict], *, n_classes: int = 71, out_path: Optional[str] = None, device="cuda",
                  max_new_tokens: int = 100) -> Dict[str, Any]:
    """Returns {'miou', 'miou_noBg', 'f_score', 'f_score_noBg', 'records'} (class means as the reference prints them)."""
    iou_pc, f_pc, cls_pc = torch.zeros(n_classes), torch.zeros(n_classes), torch.zeros(n_classes)
    records = []
    for sample in dataloader:
        sample = dict(sample)
        meta = sample.pop("batch_metadata")
        gt = sample["batch_X_modals"][0]["<mask>"]
        sample = prepare_sample(sample, device)
        res = model.generate_avs(**sample, use_cache=True, max_new_tokens=max_new_tokens)
        rec = {"instruction": meta[0].get("instruction"), "label": meta[0].get("output"),
               "predict": tokenizer.decode(res["output_ids"][0], skip_special_tokens=False)}
        masks = res.get("pred_masks")
        if masks is not None:
            i_, f_, c_, _ = M.avss_miou_fscore(masks[0].detach().float().cpu().unsqueeze(0), gt.cpu(), T=1)
            iou_pc, f_pc, cls_pc = iou_pc + i_, f_pc + f_, cls_pc + c_
            seen = (cls_pc != 0).sum()
            rec.update(iou=float(torch.nan_to_num(iou_pc / cls_pc, nan=0.0).sum() / seen),
                       fscore=float(torch.nan_to_num(f_pc / cls_pc, nan=0.0).sum() / seen))
        records.append(rec)
        write_jsonl(out_path, rec)
    miou = torch.nan_to_num(iou_pc / cls_pc, nan=0.0)
    fsc = torch.nan_to_num(f_pc / cls_pc, nan=0.0)
    return {"miou": float(miou.mean()), "miou_noBg": float(miou[:-1].mean()), "f_score": float(fsc.mean()),
            "f_score_noBg": float(fsc[:-1].mean()), "records": records}
