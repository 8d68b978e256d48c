"""Mirror of the reference's CLI surface (configs/unified_config.py:7-124): the four dataclasses
`transformers.HfArgumentParser` is built from in scripts/quick_start.py:455-456.  Field names, types and defaults are
the contract (the quick-start shell script passes them as flags); the B200 backend reads the same fields.

Grouped here by what consumes them rather than in the reference's order; un-annotated class attributes
(`select_layer_list`, `image_size` on DataArguments) are deliberately NOT dataclass fields, as in the reference, so
they are not exposed as CLI flags.
"""
from dataclasses import dataclass, field
from typing import Optional

import transformers


def _f(default, help_text=None):
    return field(default=default, metadata={"help": help_text} if help_text else None)


@dataclass
class ModelArguments:
    # decoder
    model_name_or_path: Optional[str] = _f("/data/users/henghui_du/pretrain/video-llama2/Mistral-7B-Instruct-v0.2")
    freeze_backbone: bool = _f(True, "Whether to freeze the LLM backbone.")
    llm_name: str = _f("qwen")
    # CLIP ViT tower + visual Q-Former
    vit_ckpt_path: str = _f("/group/40061/cserdu/pretrain/openai-clip-vit-large-patch14-224")
    select_layer_list = [14, 22, 23]
    select_feature: str = _f("patch")
    image_size: int = _f(224)
    patch_size: int = _f(14)
    visual_query_token_nums: int = _f(32)
    # BEATs + audio Q-Former
    BEATs_ckpt_path: str = _f("/group/40061/cserdu/pretrain/beats/BEATs_iter3_plus_AS2M_finetuned_on_AS2M_cpt2.pt")
    audio_query_token_nums: int = _f(32)
    # segmentation head (off the accelerated path; kept so the parser accepts the reference's flags)
    prompt_embed_dim: int = _f(256)
    mask_decoder_transformer_depth: int = _f(2)
    low_res_mask_size: int = _f(112)
    image_scale_nums: int = _f(2)
    token_nums_per_scale: int = _f(3)
    avs_query_num: int = _f(300)
    num_classes: int = _f(1)
    query_generator_num_layers: int = _f(2)


@dataclass
class InferenceArguments:
    ckpt_dir: str = _f("")
    avs_ckpt_dir: str = _f("")
    avss_ckpt_dir: str = _f("")
    test_name: str = _f("test")
    device: str = _f("cuda:0")


@dataclass
class DataArguments:
    video_frame_nums: int = _f(8)
    image_size = ModelArguments.image_size
    # pre-training tasks
    image_caption_task: bool = _f(False)
    video_caption_task: bool = _f(False)
    audio_caption_task: bool = _f(False)
    segmentation_task: bool = _f(False)
    # fine-tuning / evaluation tasks
    avqa_task: bool = _f(False)
    ave_task: bool = _f(False)
    avvp_task: bool = _f(False)
    arig_task: bool = _f(False)
    ms3_task: bool = _f(False)
    s4_task: bool = _f(False)
    avss_task: bool = _f(False)
    avcap_task: bool = _f(False)
    ref_avs_task: bool = _f(False)
    multi_frames: bool = _f(False)
    next_qa_task: bool = _f(False)
    aok_vqa_task: bool = _f(False)


@dataclass
class TrainingArguments(transformers.TrainingArguments):
    optim: str = _f("adamw_torch")
    mm_projector_lr: Optional[float] = None
    freeze_mm_mlp_adapter: bool = _f(False)
    remove_unused_columns: bool = _f(False)
    cache_dir: Optional[str] = _f(None)
    group_by_modality_length: bool = _f(False)
    model_max_length: int = _f(512, "Maximum sequence length. Sequences will be right padded (and possibly truncated).")
    double_quant: bool = _f(True, "Compress the quantization statistics through double quantization.")
    quant_type: str = _f("nf4", "Quantization data type to use. Should be one of `fp4` or `nf4`.")
    bits: int = _f(32, "How many bits to use.")
    lora_enable: bool = False
    lora_r: int = 8
    lora_alpha: int = 16
    lora_dropout: float = 0.05
    lora_weight_path: str = ""
    lora_bias: str = "none"
    ce_loss_weight: float = _f(1.0)
    dice_loss_weight: float = _f(0.5)
    bce_loss_weight: float = _f(2.0)
    audio_branch: bool = _f(False)
    visual_branch: bool = _f(False)
    seg_branch: bool = _f(False)
    save_modules: str = _f("vl_projector,al_projector,lora")
    exp_desc: str = _f("exp")
    use_process: bool = _f(True)
    use_hyper_lora: bool = _f(True)
    unifed_finetune_ckpt_path: str = _f("")
