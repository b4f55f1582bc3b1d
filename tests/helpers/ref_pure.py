"""Pure-Python pieces of the UNMODIFIED reference (baseline/_ref) evaluated on a grid of configurations → JSON (test helper, run as a script).
LR / WD schedules, the micro-batch ramp-up calculator, the rank generator behind parallel_state."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))
sys.path = [p for p in sys.path if os.path.abspath(p or ".") != REPO]

SCHEDULES = [
    dict(init_lr=0.0, max_lr=1e-3, min_lr=1e-5, lr_warmup_steps=10, lr_decay_steps=100, lr_decay_style="linear", start_wd=0.01, end_wd=0.1, wd_incr_steps=100, wd_incr_style="linear"),
    dict(init_lr=1e-5, max_lr=3e-4, min_lr=3e-5, lr_warmup_steps=5, lr_decay_steps=60, lr_decay_style="cosine", start_wd=0.0, end_wd=0.1, wd_incr_steps=80, wd_incr_style="cosine"),
    dict(init_lr=0.0, max_lr=1e-3, min_lr=1e-4, lr_warmup_steps=8, lr_decay_steps=100, lr_decay_style="inverse-square-root", start_wd=0.1, end_wd=0.1, wd_incr_steps=100, wd_incr_style="constant"),
    dict(init_lr=0.0, max_lr=1e-3, min_lr=1e-4, lr_warmup_steps=0, lr_decay_steps=100, lr_decay_style="constant", start_wd=0.1, end_wd=0.1, wd_incr_steps=100, wd_incr_style="constant"),
    dict(init_lr=0.0, max_lr=1e-3, min_lr=1e-5, lr_warmup_steps=10, lr_decay_steps=100, lr_decay_style="WSD", start_wd=0.1, end_wd=0.1, wd_incr_steps=100, wd_incr_style="constant",
         wsd_decay_steps=20, lr_wsd_decay_style="exponential"),
    dict(init_lr=0.0, max_lr=1e-3, min_lr=1e-5, lr_warmup_steps=10, lr_decay_steps=100, lr_decay_style="WSD", start_wd=0.1, end_wd=0.1, wd_incr_steps=100, wd_incr_style="constant",
         wsd_decay_steps=30, lr_wsd_decay_style="cosine"),
]
RAMPS = [dict(step_batch_size_schedule="0:8 64:16 128:32, 200:40", global_batch_size=None, micro_batch_size=2, data_parallel_size=2),
         dict(global_batch_size=32, micro_batch_size=4, data_parallel_size=2),
         dict(step_batch_size_schedule="0:4 1K:8 2K:16", seq_length=16, global_batch_size=None, micro_batch_size=1, data_parallel_size=4),
         dict(global_batch_size=36, micro_batch_size=4, data_parallel_size=2, decrease_batch_size_if_needed=True)]
GRIDS = [dict(tp=2, ep=1, dp=2, pp=2, cp=1, order="tp-cp-ep-dp-pp"), dict(tp=2, ep=1, dp=1, pp=1, cp=4, order="tp-cp-ep-dp-pp"), dict(tp=1, ep=4, dp=4, pp=2, cp=1, order="tp-cp-ep-dp-pp"),
         dict(tp=2, ep=2, dp=4, pp=1, cp=1, order="tp-ep-dp-pp-cp"), dict(tp=4, ep=1, dp=2, pp=2, cp=2, order="tp-pp-dp-cp-ep", rank_offset=3)]
TOKENS = ["tp", "pp", "dp", "cp", "ep", "tp-pp", "tp-dp", "dp-cp", "tp-dp-cp", "tp-ep", "tp-ep-pp", "tp-cp"]


class _Opt:
    def __init__(self):
        self.param_groups = [{"lr": 0.0, "weight_decay": 0.0, "lr_mult": 1.0, "wd_mult": 1.0}, {"lr": 0.0, "weight_decay": 0.0, "wd_mult": 0.5, "max_lr": 5e-4, "min_lr": 5e-5, "start_wd": 0.2, "end_wd": 0.2}]


def main():
    out = {"sched": [], "ramp": [], "ranks": []}
    from megatron.core.optimizer_param_scheduler import OptimizerParamScheduler

    for kw in SCHEDULES:
        opt = _Opt()
        s = OptimizerParamScheduler(opt, **kw)
        rows = []
        for _ in range(110):
            s.step(increment=1)
            rows.append([[g["lr"], g["weight_decay"]] for g in opt.param_groups])
        out["sched"].append(rows)
    from megatron.core.num_microbatches_calculator import destroy_num_microbatches_calculator, get_current_global_batch_size, get_num_microbatches, init_num_microbatches_calculator, update_num_microbatches

    for kw in RAMPS:
        init_num_microbatches_calculator(rank=0, **kw)
        rows = []
        for consumed in range(0, 320, 8):
            update_num_microbatches(consumed, consistency_check=False)
            rows.append([get_num_microbatches(), get_current_global_batch_size()])
        out["ramp"].append(rows)
        destroy_num_microbatches_calculator()
    from megatron.core.parallel_state import RankGenerator

    for kw in GRIDS:
        g = RankGenerator(**kw)
        res = {}
        for tok in TOKENS:
            try:
                res[tok] = g.get_ranks(tok)
            except Exception as e:          # tokens that the order / sizes make meaningless
                res[tok] = f"ERR {type(e).__name__}"
        out["ranks"].append(res)
    print("JSON:" + json.dumps(out))


if __name__ == "__main__":
    main()
