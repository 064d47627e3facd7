"""Checkpoint compatibility with the reference's persistence helpers.

The reference stores ``saves/<Name>-<tag>/checkpoint.pth`` dicts with the keys written by ``save_model``
(server.py:36-84): ``last_epoch, loss_record, loss_metric, last_target_name, total_training_iters, arch,
last_batch_size, model_state_dict, optimizer_state_dict, scheduler_state_dict, config``.  These helpers read and
write that schema with the B200 ``GRUUNet2`` (same state_dict keys and ``parameters()`` order, so the stored
optimizer state lines up), and mirror the loader heuristics of ``load_denoising_model_pytorch`` (app3.py:46-119).
"""
from __future__ import annotations

import datetime
import os
import uuid
from typing import Any, Mapping, Optional

import torch

from .gruunet2 import GRUUNet2

_CTOR_KEYS = ("num_compressed_bins", "in_size", "hidden_sizes", "kernel_sizes", "strides", "paddings", "num_gaussians")
_REQUIRED = _CTOR_KEYS[:-1]


def _pick_config(ckpt: Any, fallback: Optional[Mapping]) -> Optional[Mapping]:
    for key in ("hparams", "config"):  # app3.py:62-65
        if isinstance(ckpt, dict) and isinstance(ckpt.get(key), dict):
            return ckpt[key]
    for key in ("hparams", "config"):
        if not isinstance(ckpt, dict) and isinstance(getattr(ckpt, key, None), dict):
            return getattr(ckpt, key)
    return fallback


def _pick_state_dict(ckpt: Any) -> Optional[Mapping[str, torch.Tensor]]:
    if isinstance(ckpt, dict):
        for key in ("model_state_dict", "state_dict"):  # app3.py:67-70
            if key in ckpt:
                return ckpt[key]
        rest = {k: v for k, v in ckpt.items() if k not in ("hparams", "config", "last_epoch")}
        if rest and all(isinstance(v, torch.Tensor) for v in rest.values()):
            return rest
        return None
    if hasattr(ckpt, "state_dict") and callable(ckpt.state_dict):
        return ckpt.state_dict()
    return None


def _load_checkpoint_file(path: str, trust_pickle: bool):
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        if not trust_pickle:
            raise
        return torch.load(path, map_location="cpu", weights_only=False)


def load_denoising_model(path: str, fallback_config: Optional[Mapping] = None, device: Optional[torch.device] = None,
                         strict_errors: bool = False, model_class=GRUUNet2, trust_pickle: bool = False):
    """Counterpart of ``load_denoising_model_pytorch`` (app3.py:46-119): returns ``(model.eval() on device, device)`` or
    ``(None, None)`` when anything is missing or fails -- unless ``strict_errors`` asks for the exception instead.
    The file is read with ``weights_only=True`` (tensors, containers, numbers, strings: all the reference's schema needs);
    ``trust_pickle=True`` allows the full unpickler for checkpoints from a trusted source that carry other objects."""
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    try:
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        ckpt = _load_checkpoint_file(path, trust_pickle)
        sd = _pick_state_dict(ckpt)
        if sd is None:
            raise KeyError("no state dict in checkpoint")
        cfg = _pick_config(ckpt, fallback_config)
        if cfg is None:
            raise KeyError("no model config in checkpoint and no fallback given")
        kwargs = {k: cfg[k] for k in _CTOR_KEYS if k in cfg}
        missing = [k for k in _REQUIRED if k not in kwargs]
        if missing:
            raise KeyError(f"config lacks {missing}")
        model = model_class(**kwargs)
        model.load_state_dict(sd)
        model.eval()
        return model.to(dev), dev
    except Exception:
        if strict_errors:
            raise
        return None, None


def save_checkpoint(name: str, model: GRUUNet2, optimizer=None, scheduler=None, arch: Optional[str] = None, last_epoch=None,
                    loss_record=None, loss_metric=None, total_training_iters=None, last_target_name=None, last_batch_size=None,
                    tag: Optional[str] = None, allow_overwrite: bool = False, prefix: str = "saves", tag_uuid: bool = True,
                    or_tag_date: bool = True, last_dataset_name=None) -> str:
    """Write ``<prefix>/<name>-<tag>/checkpoint.pth`` with the key set of ``save_model`` (server.py:36-84); returns the path.
    Accepts the reference's keywords (``tag_uuid`` / ``or_tag_date`` / ``last_dataset_name``); ``tag="uuid" | "date" | "none"``
    is a shorthand that overrides the two booleans.  (Like the reference, ``last_dataset_name`` is accepted and not stored.)"""
    if tag is not None:
        tag_uuid, or_tag_date = tag == "uuid", tag == "date"
    if tag_uuid:
        name = f"{name}-{uuid.uuid4().hex[:6]}"
    elif or_tag_date:
        name = f"{name}-{datetime.datetime.now().strftime('%y%m%d')}"
    folder = os.path.join(prefix, name)
    if os.path.exists(folder) and not allow_overwrite:
        raise FileExistsError("File/dir already exists")
    os.makedirs(folder, exist_ok=True)
    payload = {
        "last_epoch": last_epoch,
        "loss_record": loss_record if loss_record is not None else dict(),
        "loss_metric": loss_metric,
        "last_target_name": last_target_name,
        "total_training_iters": total_training_iters,
        "last_batch_size": last_batch_size,
        "model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None,
        "scheduler_state_dict": scheduler.state_dict() if scheduler is not None else None,
        "config": model.get_config(),
        "arch": arch if arch is not None else type(model).__name__,
    }
    path = os.path.join(folder, "checkpoint.pth")
    torch.save(payload, path)
    return path


class TrainingContext:
    """``TrainingContext`` of server.py:86-142 for the B200 model: wraps the module with AdamW + ExponentialLR(0.9) and restores
    both from a checkpoint folder.  In ``train()`` mode the wrapped module's forward is differentiable (fp32 backward kernels,
    csrc/cell.cu), so ``loss.backward(); ctx.optim.step()`` works as it does with the reference module; the training LOOP itself
    (data, schedule) is the caller's, as in the reference."""

    def __init__(self, cls=GRUUNet2, *args, device: Optional[torch.device] = None, **kwargs):
        self.device = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.inner = cls(*args, **kwargs).to(self.device)
        self.name = cls.__name__
        self.optim = torch.optim.AdamW(self.inner.parameters())
        self.sched = torch.optim.lr_scheduler.ExponentialLR(self.optim, gamma=0.9)
        self.num_parameters = sum(p.numel() for p in self.inner.parameters())
        self.train_loss_record, self.test_loss_record = dict(), dict()
        self.total_iters, self.batch_size, self.best_eval_loss, self.training = 0, 64, 999, True

    def __call__(self, *args, **kwargs):
        return self.inner(*args, **kwargs)

    def __getattr__(self, item):
        inner = self.__dict__.get("inner")
        if inner is None:  # copy / pickle probe the instance before __init__ ran
            raise AttributeError(item)
        return getattr(inner, item)

    def save(self, prefix: str = "saves") -> str:
        return save_checkpoint(self.name, self.inner, optimizer=self.optim, scheduler=self.sched,
                               loss_record={"train": self.train_loss_record, "test": self.test_loss_record},
                               total_training_iters=self.total_iters, last_batch_size=self.batch_size,
                               loss_metric={"train": "MSE", "test": "MAE"}, prefix=prefix)

    @classmethod
    def load(cls, name: str, class_=GRUUNet2, prefix: str = "saves", training: bool = False, device: Optional[torch.device] = None,
             trust_pickle: bool = False):
        ckpt = _load_checkpoint_file(os.path.join(prefix, name, "checkpoint.pth"), trust_pickle)
        self = cls(class_, device=device, **ckpt["config"])
        self.inner.load_state_dict(ckpt["model_state_dict"])
        if ckpt.get("optimizer_state_dict") is not None:
            self.optim.load_state_dict(ckpt["optimizer_state_dict"])
        if ckpt.get("scheduler_state_dict") is not None:
            self.sched.load_state_dict(ckpt["scheduler_state_dict"])
        self.total_iters = ckpt.get("total_training_iters")
        self.batch_size = ckpt.get("last_batch_size")
        rec = ckpt.get("loss_record") or {}
        self.train_loss_record, self.test_loss_record = rec.get("train", {}), rec.get("test", {})
        self.best_eval_loss = min(self.test_loss_record.values()) if self.test_loss_record else None
        self.training = training
        return self
