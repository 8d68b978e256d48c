"""Batched-evaluation host code: the task loops of scripts/finetune/inference_hyper_lora.py and the metrics
of utils/avvp_eval_metrics.py, utils/avss_utils.py, utils/ciou.py, restated for this package."""
