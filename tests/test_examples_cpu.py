"""Every script under examples/ runs end to end in its TINY (CPU smoke) mode: the launch recipes cannot rot away from the entry points and flags they use."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT = [29800]

SHELL = [
    ("gpt3/train_gpt3_6.7b_tp4_pp2.sh", [], 1),
    ("llama/train_llama3_8b_tp8.sh", [], 1),
    ("llama/train_llama3_70b_tp8.sh", [], 1),
    ("mixtral/train_mixtral_8x7b_ep8.sh", [], 1),
    ("deepseek_mla/train_mla_moe.sh", [], 1),
    ("megatron_fsdp/train_llama3_8b_fsdp.sh", [], 2),
    ("bert/train_bert_340m.sh", [], 1),
    ("t5/train_t5_220m.sh", [], 1),
    ("mamba/train_hybrid_mamba2.sh", [], 1),
    ("multimodal/train_llava_style.sh", [], 1),
    ("long_context/train_llama_cp2_sliding_window.sh", ["--context-parallel-size", "2", "--cp-comm-type", "p2p"], 2),
    ("long_context/train_llama_cp2_sliding_window.sh", ["--window-size", "15", "0"], 1),
]


def _run(cmd, env_extra, timeout=600):
    PORT[0] += 1
    env = dict(os.environ, TINY="1", CUDA_VISIBLE_DEVICES="", MASTER_PORT=str(PORT[0]), WANDB_MODE="offline", **env_extra)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-2500:]
    return r.stdout


@pytest.mark.parametrize("script,extra,nproc", SHELL, ids=[f"{s[0]}{'+' + s[1][0] if s[1] else ''}" for s in SHELL])
def test_training_example_runs_in_tiny_mode(script, extra, nproc):
    out = _run(["bash", os.path.join(ROOT, "examples", script)] + extra, {"NPROC": str(nproc)})
    assert "iteration        2/       2" in out and "lm loss" in out


def test_rl_example(tmp_path):
    out = _run(["bash", os.path.join(ROOT, "examples", "rl/train_grpo.sh")], {"RL_PROFILE_DIR": str(tmp_path)})
    assert "iter   2 |" in out and "rl profile:" in out and (tmp_path / "rl_profile_rank0.jsonl").exists()


def test_python_examples(tmp_path):
    out = _run([sys.executable, os.path.join(ROOT, "examples/export/export_hf_and_trtllm.py"), "--out", str(tmp_path / "export")], {})
    assert "max abs difference 0.00e+00" in out and "rank0.safetensors" in out
    out = _run([sys.executable, os.path.join(ROOT, "examples/post_training/distill_and_quantize.py"), "--iters", "2"], {})
    assert "quantised 4 linear layers" in out and "finite = True" in out
    out = _run([sys.executable, os.path.join(ROOT, "examples/inference/zmq_data_parallel_serving.py")], {})
    assert "'served': [3, 3]" in out
    out = _run([sys.executable, os.path.join(ROOT, "examples/run_simple_mcore_train_loop.py")], {}) if os.path.exists(os.path.join(ROOT, "examples/run_simple_mcore_train_loop.py")) else ""
