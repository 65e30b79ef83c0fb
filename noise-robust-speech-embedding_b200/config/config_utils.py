"""YAML + argparse configuration with the reference's surface (ref:config/config_utils.py:6-66): a plain nested
dict with sections model / training / data / logging / emotion, CLI overrides --device --batch_size --epochs --lr.
New optional keys (all default to the reference behaviour): ``model.pooling`` ("mean"), ``model.frontend`` ("b200"),
``training.gpu_mix`` (true)."""
from __future__ import annotations

import argparse
from typing import Any, Dict, Optional, Sequence

import yaml


def load_config(path: str) -> Dict[str, Any]:
    with open(path, "r") as f:
        return yaml.safe_load(f)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Noise-robust speech embedding (B200 hot path)")
    p.add_argument("--config", type=str, default="config/default.yaml")
    p.add_argument("--device", type=str, default=None)
    p.add_argument("--batch_size", type=int, default=None)
    p.add_argument("--epochs", type=int, default=None)
    p.add_argument("--lr", type=float, default=None)
    p.add_argument("--task", type=str, default=None)
    return p


def get_config(argv: Optional[Sequence[str]] = None) -> Dict[str, Any]:
    args = build_parser().parse_args(argv)
    config = load_config(args.config)
    config["device"] = args.device or config.get("device", "cuda:0")
    training = config.setdefault("training", {})
    if args.batch_size is not None:
        training["batch_size"] = args.batch_size
    if args.epochs is not None:
        training["num_epochs"] = args.epochs
    if args.lr is not None:
        training["learning_rate"] = args.lr
    if args.task is not None:
        config["task"] = args.task
    return config
