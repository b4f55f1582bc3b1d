#!/usr/bin/env python
"""Hybrid (Mamba-2 / attention / MLP / MoE pattern) pre-training — the reference's ``pretrain_hybrid.py`` name for what ``pretrain_mamba.py`` runs."""
import runpy
import os

if __name__ == "__main__":
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pretrain_mamba.py"), run_name="__main__")
