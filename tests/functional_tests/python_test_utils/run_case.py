"""Run one functional test case and compare with its golden values (reference ``tests/functional_tests/python_test_utils/test_pretraining_*_pipeline.py``)."""
from __future__ import annotations

import argparse
import json
import os
import shutil
import socket
import subprocess
import sys
import tempfile
from typing import Dict, List

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _argv(model_args: dict) -> List[str]:
    out = []
    for k, v in model_args.items():
        if v is True:
            out.append(k)
        elif v is False or v is None:
            continue
        else:
            out += [k, str(v)]
    return out


def launch(cfg: dict, extra: List[str], tb_dir: str, script: str = "pretrain_gpt.py", timeout: int = 900) -> str:
    nproc = int(cfg.get("NPROC", 1))
    env = dict(os.environ, **{k: str(v) for k, v in (cfg.get("ENV_VARS") or {}).items()})
    args = _argv(cfg["MODEL_ARGS"]) + ["--tensorboard-dir", tb_dir] + extra
    if nproc == 1:
        env.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_port()), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
        cmd = [sys.executable, os.path.join(ROOT, script)] + args
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", str(_port()),
               os.path.join(ROOT, script)] + args
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"training failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}")
    return r.stdout


def read_scalars(tb_dir: str) -> Dict[str, Dict[int, float]]:
    path = os.path.join(tb_dir, "scalars.jsonl")
    out: Dict[str, Dict[int, float]] = {}
    if os.path.exists(path):
        with open(path) as f:
            for line in f:
                r = json.loads(line)
                if "value" in r:
                    out.setdefault(r["tag"], {})[int(r["step"])] = r["value"]
        return out
    from tensorboard.backend.event_processing import event_accumulator      # real TensorBoard event files

    ea = event_accumulator.EventAccumulator(tb_dir, size_guidance={"scalars": 0})
    ea.Reload()
    for tag in ea.Tags()["scalars"]:
        out[tag] = {e.step: e.value for e in ea.Scalars(tag)}
    return out


def run_regular(case_dir: str, cfg: dict, update: bool, platform: str) -> Dict[str, Dict[int, float]]:
    with tempfile.TemporaryDirectory() as tmp:
        launch(cfg, ["--train-iters", str(cfg["TRAIN_ITERS"])], os.path.join(tmp, "tb"))
        got = read_scalars(os.path.join(tmp, "tb"))
    golden_path = os.path.join(case_dir, f"golden_values_{platform}.json")
    keep = {k: {str(s): v for s, v in got[k].items()} for k in ("lm loss",) if k in got}
    if update or not os.path.exists(golden_path):
        with open(golden_path, "w") as f:
            json.dump(keep, f, indent=1, sort_keys=True)
        return got
    golden = json.load(open(golden_path))
    for tag, series in golden.items():
        for step, ref in series.items():
            val = got[tag][int(step)]
            if abs(val - ref) > 1e-4 * max(abs(ref), 1.0):
                raise AssertionError(f"{tag} at iteration {step}: {val} vs golden {ref}")
    return got


def run_resume(case_dir: str, cfg: dict) -> None:
    n = int(cfg["TRAIN_ITERS"])
    half = n // 2
    with tempfile.TemporaryDirectory() as tmp:
        full_tb = os.path.join(tmp, "tb_full")
        launch(cfg, ["--train-iters", str(n)], full_tb)
        ck = os.path.join(tmp, "ckpt")
        common = ["--train-iters", str(n), "--save", ck, "--load", ck, "--save-interval", str(half)]
        launch(cfg, common + ["--exit-interval", str(half)], os.path.join(tmp, "tb_a"))
        launch(cfg, common, os.path.join(tmp, "tb_b"))
        full, tail = read_scalars(full_tb)["lm loss"], read_scalars(os.path.join(tmp, "tb_b"))["lm loss"]
    for step in range(half + 1, n + 1):
        if full[step] != tail[step]:
            raise AssertionError(f"resume mismatch at iteration {step}: {tail[step]} vs uninterrupted {full[step]}")


def run_reshard(case_dir: str, cfg: dict, rtol: float = 2e-3) -> None:
    """Reference ``*_reshard_*`` cases: save under layout A, resume under layout B — the loss curve must continue as if nothing happened (to fp tolerance: the
    reduction orders of the two layouts differ)."""
    import copy

    n = int(cfg["TRAIN_ITERS"])
    half = n // 2
    a, b = copy.deepcopy(cfg), copy.deepcopy(cfg)
    a["MODEL_ARGS"].update(cfg.get("LAYOUT_A") or {})
    b["MODEL_ARGS"].update(cfg.get("LAYOUT_B") or {})
    with tempfile.TemporaryDirectory() as tmp:
        launch(a, ["--train-iters", str(n)], os.path.join(tmp, "tb_full"))
        ck = os.path.join(tmp, "ckpt")
        common = ["--train-iters", str(n), "--save", ck, "--load", ck, "--save-interval", str(half)]
        launch(a, common + ["--exit-interval", str(half)], os.path.join(tmp, "tb_a"))
        launch(b, common, os.path.join(tmp, "tb_b"))
        full, tail = read_scalars(os.path.join(tmp, "tb_full"))["lm loss"], read_scalars(os.path.join(tmp, "tb_b"))["lm loss"]
    assert min(tail) == half + 1, f"resumed run did not start at iteration {half + 1}: {sorted(tail)}"
    for step in range(half + 1, n + 1):
        if abs(full[step] - tail[step]) > rtol * abs(full[step]):
            raise AssertionError(f"reshard-resume mismatch at iteration {step}: {tail[step]} vs {full[step]}")


def run_case(case_dir: str, update_golden: bool = False, platform: str = "cpu") -> None:
    cfg = yaml.safe_load(open(os.path.join(case_dir, "model_config.yaml")))
    types = cfg.get("TEST_TYPE", ["regular"])
    if "regular" in types:
        run_regular(case_dir, cfg, update_golden, platform)
    if "ckpt-resume" in types and not update_golden:
        run_resume(case_dir, cfg)
    if "ckpt-reshard" in types and not update_golden:
        run_reshard(case_dir, cfg)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case_dir")
    ap.add_argument("--update-golden", action="store_true")
    ap.add_argument("--platform", default="cpu")
    a = ap.parse_args()
    run_case(a.case_dir, a.update_golden, a.platform)
    print("PASSED", a.case_dir)
