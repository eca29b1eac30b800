"""VisionModelSAERunnerConfig -- the SAE run description (reference sae/config.py:287-663).

Same fields, defaults and derived properties as the reference dataclass so notebooks
(demos/2_Train_SAE.ipynb cells 0-6) construct it unchanged.  Differences, all additive:
  * ``dtype_mapping`` knows ``bfloat16`` (the reference raises KeyError for it, config.py:14-45);
  * ``num_patch`` works (the reference forgets ``import math``);
  * ``CacheActivationsRunnerConfig.__post_init__`` no longer calls a non-existent base class.
"""
from __future__ import annotations

import inspect
import json
import logging
import math
import os
from dataclasses import asdict, dataclass, field, fields
from typing import Any, Literal, Optional

import torch

from vit_prisma.configs.HookedViTConfig import HookedViTConfig

_BASE_DTYPES = {
    "float32": torch.float32, "float": torch.float32, "float64": torch.float64, "double": torch.float64,
    "float16": torch.float16, "half": torch.float16, "bfloat16": torch.bfloat16, "bf16": torch.bfloat16,
    "int64": torch.int64, "long": torch.int64, "int32": torch.int32, "int": torch.int32,
    "int16": torch.int16, "short": torch.int16, "int8": torch.int8, "uint8": torch.uint8, "bool": torch.bool,
}
dtype_mapping = {**_BASE_DTYPES, **{f"torch.{k}": v for k, v in _BASE_DTYPES.items()}}


@dataclass
class VisionModelSAERunnerConfig:
    # ---- what the SAE is attached to -------------------------------------------------
    model_class_name: str = "HookedViT"
    model_name: str = "open-clip:laion/CLIP-ViT-B-32-DataComp.XL-s13B-b90K"
    vit_model_cfg: Optional[HookedViTConfig] = None
    model_path: str = None
    hook_point_layer: int = 9
    layer_subtype: str = "ln2.hook_normalized"
    hook_point_head_index: Optional[int] = None
    context_size: int = 50
    use_cached_activations: bool = False
    use_patches_only: bool = False
    cached_activations_path: Optional[str] = None   # default: activations/{dataset}/{model}/{hook}[_{head}]
    image_size: int = 224
    architecture: Literal["standard", "gated", "jumprelu"] = "standard"

    # ---- SAE shape / init -------------------------------------------------------------
    b_dec_init_method: str = "geometric_median"
    expansion_factor: int = 16
    from_pretrained_path: Optional[str] = None

    # ---- transcoder -------------------------------------------------------------------
    is_transcoder: bool = False
    transcoder_with_skip_connection: bool = True
    out_hook_point_layer: int = 9
    layer_out_subtype: str = "hook_mlp_out"
    d_out: int = 768

    # ---- device / dtype (string-backed, see the properties) ----------------------------
    _device: str = "cuda"
    seed: int = 42
    _dtype: str = "float32"

    d_in: int = 768
    activation_fn_str: str = "topk"          # "relu" | "topk" | "tanh-relu"
    activation_fn_kwargs: dict = field(default_factory=dict)
    cls_token_only: bool = False

    max_grad_norm: float = 1.0               # None / 0 turns clipping off
    initialization_method: str = "independent"   # or "encoder_transpose_decoder"
    normalize_activations: str = "layer_norm"

    # ---- training ---------------------------------------------------------------------
    is_training = True

    n_batches_in_buffer: int = 20
    store_batch_size: int = 32
    num_workers: int = 16
    num_epochs: int = 1
    verbose: bool = False

    l1_coefficient: float = 0.0002
    lp_norm: float = 1
    lr: float = 0.001
    lr_scheduler_name: str = "cosineannealingwarmup"
    lr_warm_up_steps: int = 500
    train_batch_size: int = 1024 * 4

    min_l0 = None
    min_explained_variance = None

    dataset_name: str = "imgnet"
    dataset_path: str = "/network/scratch/s/sonia.joseph/datasets/kaggle_datasets"
    dataset_train_path: str = "/network/scratch/s/sonia.joseph/datasets/kaggle_datasets/ILSVRC/Data/CLS-LOC/train"
    dataset_val_path: str = "/network/scratch/s/sonia.joseph/datasets/kaggle_datasets/ILSVRC/Data/CLS-LOC/val"

    # ---- dead-feature handling ----------------------------------------------------------
    use_ghost_grads: bool = False
    feature_sampling_window: int = 1000
    dead_feature_window: int = 5000
    dead_feature_threshold: float = 1e-8

    # ---- logging / checkpoints -------------------------------------------------------
    log_to_wandb: bool = True
    wandb_project: str = "tinyclip_sae_16_hyperparam_sweep_lr"
    wandb_entity: Optional[str] = None
    wandb_log_frequency: int = 10
    n_validation_runs: int = 0
    n_checkpoints: int = 10
    checkpoint_path: str = "/network/scratch/p/praneet.suresh/open_clip_celeba_checkpoints/"

    # ------------------------------------------------------------------ string-backed views
    @property
    def device(self):
        return torch.device(self._device) if isinstance(self._device, str) else self._device

    @device.setter
    def device(self, value):
        self._device = value

    @property
    def dtype(self):
        return dtype_mapping[self._dtype]

    @dtype.setter
    def dtype(self, value):
        self._dtype = value

    # ------------------------------------------------------------------ derived
    @property
    def hook_point(self) -> str:
        return f"blocks.{self.hook_point_layer}.{self.layer_subtype}"

    @hook_point.setter
    def hook_point(self, value):
        self._custom_hook_point = value

    @property
    def out_hook_point(self) -> str:
        return f"blocks.{self.out_hook_point_layer}.{self.layer_out_subtype}"

    @property
    def _tokens_per_image(self) -> int:
        if self.cls_token_only:
            return 1
        return self.context_size - 1 if self.use_patches_only else self.context_size

    @property
    def tokens_per_buffer(self) -> int:
        return self.train_batch_size * self._tokens_per_image * self.n_batches_in_buffer

    @property
    def total_training_images(self) -> int:
        return int(1_300_000 * self.num_epochs)

    @property
    def total_training_tokens(self) -> int:
        return self.total_training_images * self._tokens_per_image

    @property
    def total_training_steps(self) -> int:
        return self.total_training_tokens // self.train_batch_size

    @property
    def d_sae(self) -> int:
        return self.d_in * self.expansion_factor

    @property
    def num_patch(self) -> int:
        return int(math.sqrt(self.context_size - 1))

    # ------------------------------------------------------------------ validation
    def __post_init__(self):
        if self.b_dec_init_method not in ("geometric_median", "mean", "zeros"):
            raise ValueError(f"b_dec_init_method must be geometric_median, mean, or zeros. Got {self.b_dec_init_method}")
        if self.b_dec_init_method == "zeros":
            logging.warning("Warning: We are initializing b_dec to zeros. This is probably not what you want.")
        if self.cls_token_only and self.use_patches_only:
            raise ValueError("cls_token_only and use_patches_only are exclusive.")
        if self.cached_activations_path is None:
            self.cached_activations_path = (
                f"activations/{self.dataset_path.replace('/', '_')}/{self.model_name.replace('/', '_')}/{self.hook_point}")
            if self.hook_point_head_index is not None:
                self.cached_activations_path += f"_{self.hook_point_head_index}"
        if os.getenv("EVAL_MODE", "false").lower() in {"true", "1"}:
            self.is_training = False
            logging.info("Evaluation mode detected via environment variable; setting is_training to False.")
        logging.info(f"Total training steps: {self.total_training_steps}; expansion factor: {self.expansion_factor}; "
                     f"tokens per buffer (M): {self.store_batch_size * self.context_size * self.n_batches_in_buffer / 1e6}")
        if self.use_ghost_grads:
            logging.info("Using Ghost Grads.")

    def is_property(self, attr_name: str) -> bool:
        return isinstance(getattr(self.__class__, attr_name, None), property)

    # ------------------------------------------------------------------ persistence
    def save_config(self, path: str) -> None:
        def plain(obj):
            if inspect.isdatadescriptor(obj):
                return None
            if isinstance(obj, (list, tuple)):
                return [plain(x) for x in obj]
            if isinstance(obj, dict):
                return {k: plain(v) for k, v in obj.items() if not self.is_property(k)}
            if isinstance(obj, (torch.dtype, torch.device)):
                return str(obj)
            return obj

        data = plain(asdict(self))
        data["_dtype"], data["_device"] = self._dtype, self._device
        with open(path, "w") as f:
            json.dump(data, f, indent=4)

    @classmethod
    def load_config(cls, path: str) -> "VisionModelSAERunnerConfig":
        with open(path, "r") as f:
            data = json.load(f)
        for legacy in ("total_training_images", "total_training_tokens", "d_sae"):
            if legacy in data:
                logging.warning(f"Deprecated field '{legacy}' found in config. It will be ignored.")
                del data[legacy]
        known = {f.name for f in fields(cls)}
        data = {k: v for k, v in data.items() if k in known}
        if isinstance(data.get("vit_model_cfg"), dict):
            vit = dict(data["vit_model_cfg"])
            if isinstance(vit.get("dtype"), str):
                vit["dtype"] = dtype_mapping.get(vit["dtype"], torch.float32)
            data["vit_model_cfg"] = HookedViTConfig(**vit)
        return cls(**data)

    def pretty_print(self) -> None:
        print("Configuration:")
        for f in fields(self):
            value = getattr(self, f.name)
            if isinstance(value, torch.dtype):
                value = str(value).split(".")[-1]
            elif isinstance(value, torch.device):
                value = str(value)
            print(f"  {f.name}: {value}")


@dataclass
class CacheActivationsRunnerConfig(VisionModelSAERunnerConfig):
    """Activation-caching run (reference :665-681, whose __post_init__ cannot run as written)."""
    shuffle_every_n_buffers: int = 10
    n_shuffles_with_last_section: int = 10
    n_shuffles_in_entire_dir: int = 10
    n_shuffles_final: int = 100

    def __post_init__(self):
        super().__post_init__()
        if self.use_cached_activations:
            raise ValueError("Use_cached_activations should be False when running cache_activations_runner")
