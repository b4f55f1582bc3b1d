"""Checkpoint lineage for the W&B sink (reference ``training/wandb_utils.py``): record which checkpoint a run saved / loaded as artifacts
(or, with the offline JSON writer, as text records)."""
from __future__ import annotations

import os

from .global_vars import get_wandb_writer


def _tracker(path: str) -> str:
    return os.path.join(path, "latest_wandb_artifact_path.txt")


def on_save_checkpoint_success(checkpoint_path: str, tracker_filename: str, save_dir: str, iteration: int) -> None:
    w = get_wandb_writer()
    if w is None:
        return
    name = f"{os.path.basename(os.path.normpath(save_dir))}-iter{iteration}"
    if hasattr(w, "Artifact"):
        art = w.Artifact(name.replace("/", "_"), type="model", metadata={"iteration": iteration})
        art.add_reference(f"file://{os.path.abspath(checkpoint_path)}", checksum=False)
        art.add_file(tracker_filename)
        w.run.log_artifact(art, aliases=[f"iter{iteration}"])
        ref = f"{w.run.entity}/{w.run.project}/{art.name}"
    else:
        w.add_text("checkpoint/saved", os.path.abspath(checkpoint_path), iteration)
        ref = name
    with open(_tracker(save_dir), "w") as f:
        f.write(ref)


def on_load_checkpoint_success(checkpoint_path: str, load_dir: str) -> None:
    w = get_wandb_writer()
    if w is None:
        return
    ref = None
    if os.path.exists(_tracker(load_dir)):
        with open(_tracker(load_dir)) as f:
            ref = f.read().strip()
    if hasattr(w, "run") and ref:
        try:
            w.run.use_artifact(ref)
        except Exception:
            pass
    elif hasattr(w, "add_text"):
        w.add_text("checkpoint/loaded", os.path.abspath(checkpoint_path))
