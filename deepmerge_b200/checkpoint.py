"""Checkpoint files of the reference: Train_SMT.py:318-340 saves a dictionary with the
keys net / optimizer / epoch / time / scales / depth / name under a time-stamped file name, and every consumer
(Train_SMT.py:164-198, ExtractFeatures.py:35-36) reads `torch.load(path)['net']`.  Same keys, same file-name pattern,
so checkpoints move between the reference and this package unchanged (e.g. a pair-MLP trained with the reference's
loop scoring edges here through deepmerge_b200.Nets.MLP)."""
from __future__ import annotations

import time as _time

import torch

KEYS = ("net", "optimizer", "epoch", "time", "scales", "depth", "name")


def model_name(epoch, real_time=None):
    """'model-Y-M-D_H-M_<epoch+1>epochs.pth' (Train_SMT.py:318-324)."""
    t = _time.localtime() if real_time is None else real_time
    return "model-{0}-{1}-{2}_{3}-{4}_{5}epochs.pth".format(t.tm_year, t.tm_mon, t.tm_mday, t.tm_hour, t.tm_min, epoch + 1)


def state(net, optimizer, epoch, elapsed):
    """The dictionary the reference saves; `elapsed` = seconds trained so far (it stores round(..., 2))."""
    return {"net": net.state_dict(),
            "optimizer": optimizer.state_dict() if optimizer is not None else None,
            "epoch": epoch,
            "time": round(float(elapsed), 2),
            "scales": getattr(net, "input_image_scales", None),
            "depth": getattr(net, "depth", None),
            "name": getattr(net, "name", type(net).__name__)}


def save(path, net, optimizer, epoch, elapsed):
    torch.save(state(net, optimizer, epoch, elapsed), path)
    return path


def load(path, net=None, optimizer=None, map_location="cpu", strict=True, trusted=False):
    """torch.load + `net.load_state_dict(ckpt['net'])` (+ optimizer) as the reference's resume code does.
    A bare state_dict (the reference's `weights` path, Train_SMT.py:182-188) is accepted as well.
    -> the checkpoint dictionary (epoch, time, ... for the caller).

    The dictionary only holds state_dicts, numbers, lists and strings, so it is read with the safe loader
    (weights_only=True: no pickle code runs).  trusted=True falls back to the full unpickler for legacy files that hold
    other objects -- only for files whose origin you trust."""
    ckpt = torch.load(path, map_location=map_location, weights_only=not trusted)
    if not (isinstance(ckpt, dict) and "net" in ckpt):
        ckpt = {"net": ckpt, "optimizer": None, "epoch": -1, "time": 0.0, "scales": None, "depth": None, "name": None}
    if net is not None:
        net.load_state_dict(ckpt["net"], strict=strict)
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    return ckpt
