"""Forward/backward schedules: no pipelining, 1F1B, interleaved 1F1B
(reference ``pipeline_parallel/schedules.py:723, 2147, 1019``).

All three share ``forward_step`` / ``backward_step``; the pipelined schedules are written
as an explicit *plan* — a list of (op, microbatch, chunk) tuples generated up front
(``build_interleaved_plan``) and then executed — so the ordering logic is a pure function
that is unit-tested on CPU without any process group.
"""
from __future__ import annotations

import contextlib
from typing import Iterator, List, Optional, Tuple, Union

import torch
from torch.autograd.variable import Variable

from .. import parallel_state as ps
from ..utils import get_attr_wrapped_model, get_model_config, get_model_type, get_pg_size
from .p2p_communication import P2PCommunicator

Shape = Union[List[int], torch.Size]


def get_forward_backward_func(pp_size: Optional[int] = None, vp_size: Optional[int] = None, schedule_pg_collection=None):
    """Pick the schedule for the current pipeline configuration.

    The returned function takes ``forward_step_func(data_iterator, model) -> (output,
    loss_func)``, ``data_iterator``, ``model`` (list of chunks for interleaving),
    ``num_microbatches``, ``seq_length``, ``micro_batch_size`` …, and returns the list
    of per-micro-batch loss dicts on the last stage.
    """
    pp = ps.get_pipeline_model_parallel_world_size() if pp_size is None else pp_size
    vp = ps.get_virtual_pipeline_model_parallel_world_size() if vp_size is None else vp_size
    if pp > 1:
        return forward_backward_pipelining_with_interleaving if vp is not None else forward_backward_pipelining_without_interleaving
    return forward_backward_no_pipelining


def deallocate_output_tensor(out, deallocate_pipeline_outputs=False):
    """Free an activation that was sent downstream: only its ``grad_fn`` is needed later."""
    if out is None or not deallocate_pipeline_outputs:
        return
    assert isinstance(out, torch.Tensor) and out._base is None, "counter-productive to free a view"
    out.data = torch.empty((1,), device=out.device, dtype=out.dtype)


def custom_backward(output, grad_output):
    """``torch.autograd.backward`` without the shape check (output was deallocated)."""
    assert output.numel() == 1, "output should be pseudo-freed in schedule, to optimize memory"
    assert isinstance(output, torch.Tensor) and isinstance(grad_output, (torch.Tensor, type(None)))
    if grad_output is None:
        assert output.numel() == 1, "implicit grad requires scalar output"
        grad_output = torch.ones_like(output, memory_format=torch.preserve_format)
    Variable._execution_engine.run_backward(
        tensors=(output,), grad_tensors=(grad_output,), keep_graph=False, create_graph=False, inputs=tuple(),
        allow_unreachable=True, accumulate_grad=True,
    )


def set_current_microbatch(model, microbatch_id):
    try:
        decoder = get_attr_wrapped_model(model, "decoder")
    except RuntimeError:
        decoder = None
    if decoder is not None and hasattr(decoder, "layers"):
        for layer in decoder.layers:
            layer.current_microbatch = microbatch_id


def forward_step_calc_loss(model, output_tensor, loss_func, config, vp_stage, collect_non_loss_data, num_microbatches, forward_data_store,
                           cp_group_size=None, is_last_stage=None):
    """Run the user loss on the last stage; returns ``(output_tensor, num_tokens)``."""
    # on the accelerator (reference: device="cuda") so the pp broadcast / dp x cp all-reduce in finalize_model_grads work over NCCL
    _dev = output_tensor.device if isinstance(output_tensor, torch.Tensor) else (output_tensor[0].device if output_tensor else torch.device("cpu"))
    num_tokens = torch.tensor(0, dtype=torch.int, device=_dev)
    if is_last_stage is None:
        is_last_stage = ps.is_pipeline_last_stage(ignore_virtual=False, vp_stage=vp_stage)
    if cp_group_size is None:
        cp_group_size = ps.get_context_parallel_world_size()
    if is_last_stage:
        if not collect_non_loss_data:
            outputs = loss_func(output_tensor)
            if len(outputs) == 3:
                output_tensor, num_tokens, loss_reduced = outputs
                if not config.calculate_per_token_loss:
                    output_tensor = output_tensor / num_tokens.clamp(min=1)
                    output_tensor = output_tensor / num_microbatches
            else:
                assert len(outputs) == 2
                output_tensor, loss_reduced = outputs
                output_tensor = output_tensor * cp_group_size / num_microbatches
            forward_data_store.append(loss_reduced)
        else:
            forward_data_store.append(loss_func(output_tensor, non_loss_data=True))
    if config.timers is not None:
        pass
    # Auxiliary losses injected through autograd "scaler" functions must see the same factor as the main loss
    # (reference schedules.py:355-397): loss_scale * cp / num_microbatches, or the bare loss scale with per-token loss.
    per_token = config.calculate_per_token_loss
    has_moe = getattr(config, "num_moe_experts", None) is not None
    has_mtp = getattr(config, "mtp_num_layers", None) is not None
    has_dsa = getattr(config, "experimental_attention_variant", None) == "dsa" or float(getattr(config, "dsa_indexer_loss_coeff", 0.0) or 0.0) > 0
    if has_moe or has_mtp or has_dsa:
        dev = output_tensor.device if isinstance(output_tensor, torch.Tensor) else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() and not config.use_cpu_initialization else torch.device("cpu"))

        def _scale(func):
            one = torch.ones(1, device=dev)
            return func(one) if func is not None else one

        cp_scale = cp_group_size if cp_group_size is not None else 1
        if has_moe:
            from ..transformer.moe.moe_utils import MoEAuxLossAutoScaler

            ls = _scale(config.grad_scale_func)
            MoEAuxLossAutoScaler.set_loss_scale(ls if per_token else ls * cp_scale / num_microbatches)
        if has_mtp:
            from ..transformer.multi_token_prediction import MTPLossAutoScaler

            ls = _scale(getattr(config, "mtp_grad_scale_func", None) or config.grad_scale_func)
            MTPLossAutoScaler.set_loss_scale(ls if per_token else ls / num_microbatches)
        if has_dsa:
            from ..transformer.experimental_attention_variant.dsa import DSAIndexerLossAutoScaler

            ls = _scale(config.grad_scale_func)
            DSAIndexerLossAutoScaler.set_loss_scale(ls if per_token else ls * cp_scale / num_microbatches)
    return output_tensor, num_tokens


def forward_step(forward_step_func, data_iterator, model, num_microbatches, input_tensor, forward_data_store, config,
                 cp_group_size=None, collect_non_loss_data=False, checkpoint_activations_microbatch=None, is_first_microbatch=False,
                 current_microbatch=None, vp_stage=None, is_last_stage=None):
    if config.timers is not None:
        config.timers("forward-compute", log_level=2).start()
    if is_first_microbatch and hasattr(model, "set_is_first_microbatch"):
        model.set_is_first_microbatch()
    if current_microbatch is not None:
        set_current_microbatch(model, current_microbatch)
    unwrap_output_tensor = False
    if not isinstance(input_tensor, list):
        input_tensor = [input_tensor]
        unwrap_output_tensor = True
    get_attr_wrapped_model(model, "set_input_tensor")(input_tensor)
    ctx = torch.autocast("cuda", dtype=config.autocast_dtype) if config.enable_autocast else contextlib.nullcontext()
    with ctx:
        if checkpoint_activations_microbatch is None:
            output_tensor, loss_func = forward_step_func(data_iterator, model)
        else:
            output_tensor, loss_func = forward_step_func(data_iterator, model, checkpoint_activations_microbatch)
    output_tensor, num_tokens = forward_step_calc_loss(
        model, output_tensor, loss_func, config, vp_stage, collect_non_loss_data, num_microbatches, forward_data_store,
        cp_group_size=cp_group_size, is_last_stage=is_last_stage,
    )
    if config.timers is not None:
        config.timers("forward-compute").stop()
    if unwrap_output_tensor:
        return output_tensor, num_tokens
    return [output_tensor], num_tokens


def backward_step(input_tensor, output_tensor, output_tensor_grad, model_type, config):
    """Backward through one stage; returns the gradient w.r.t. the stage input."""
    if config.timers is not None:
        config.timers("backward-compute", log_level=2).start()
    unwrap = False
    if not isinstance(input_tensor, list):
        input_tensor, unwrap = [input_tensor], True
    for x in input_tensor:
        if x is not None:
            x.retain_grad()
    if not isinstance(output_tensor, list):
        output_tensor = [output_tensor]
    if not isinstance(output_tensor_grad, list):
        output_tensor_grad = [output_tensor_grad]
    if output_tensor_grad[0] is None and config.grad_scale_func is not None:
        output_tensor[0] = config.grad_scale_func(output_tensor[0])
    if output_tensor[0].requires_grad:
        if config.deallocate_pipeline_outputs:
            custom_backward(output_tensor[0], output_tensor_grad[0])
        else:
            torch.autograd.backward(output_tensor[0], grad_tensors=output_tensor_grad[0])
    grads = [None if x is None else x.grad for x in input_tensor]
    if config.timers is not None:
        config.timers("backward-compute").stop()
    return grads[0] if unwrap else grads


def check_first_val_step(first_val_step, forward_only, cond):
    return (first_val_step and cond) if (first_val_step is not None and forward_only) else cond


def _finalize(config, model, total_num_tokens, forward_only, force_all_reduce=False, pg_collection=None):
    if config.finalize_model_grads_func is not None and not forward_only:
        config.finalize_model_grads_func(
            model if isinstance(model, list) else [model],
            total_num_tokens if config.calculate_per_token_loss else None,
            **({"pg_collection": pg_collection} if pg_collection is not None else {}),
        )


# ==================================================================================
# no pipelining
# ==================================================================================


def forward_backward_no_pipelining(*, forward_step_func, data_iterator: Union[Iterator, List[Iterator]], model, num_microbatches: int,
                                   seq_length: int, micro_batch_size: int, decoder_seq_length: Optional[int] = None, forward_only: bool = False,
                                   collect_non_loss_data: bool = False, first_val_step: Optional[bool] = None, adjust_tensor_shapes_fn=None,
                                   p2p_communicator=None, pg_collection=None, force_all_reduce: bool = False):
    """Gradient accumulation over ``num_microbatches``; DP grad sync only on the last one."""
    if isinstance(model, list):
        assert len(model) == 1, "non-pipeline-parallel schedule does not support model chunking"
        model = model[0]
    if isinstance(data_iterator, list):
        assert len(data_iterator) == 1
        data_iterator = data_iterator[0]
    config = get_model_config(model)
    if config.timers is not None:
        config.timers("forward-backward", log_level=1).start(barrier=config.barrier_with_L1_time)
    no_sync_func = config.no_sync_func or contextlib.nullcontext
    model_type = get_model_type(model)
    forward_data_store = []
    total_num_tokens = torch.zeros([], dtype=torch.int)
    with no_sync_func():
        for i in range(num_microbatches - 1):
            out, nt = forward_step(forward_step_func, data_iterator, model, num_microbatches, None, forward_data_store, config,
                                   collect_non_loss_data=collect_non_loss_data, is_first_microbatch=check_first_val_step(first_val_step, forward_only, i == 0),
                                   current_microbatch=i, is_last_stage=True)
            total_num_tokens = total_num_tokens.to(nt.device) + nt
            if not forward_only:
                backward_step(None, out, None, model_type, config)
    out, nt = forward_step(forward_step_func, data_iterator, model, num_microbatches, None, forward_data_store, config,
                           collect_non_loss_data=collect_non_loss_data,
                           is_first_microbatch=check_first_val_step(first_val_step, forward_only, num_microbatches == 1),
                           current_microbatch=num_microbatches - 1, is_last_stage=True)
    total_num_tokens = total_num_tokens.to(nt.device) + nt
    if not forward_only:
        backward_step(None, out, None, model_type, config)
    _finalize(config, model, total_num_tokens, forward_only, force_all_reduce, pg_collection)
    if config.timers is not None:
        config.timers("forward-backward").stop()
    return forward_data_store


# ==================================================================================
# 1F1B (non-interleaved)
# ==================================================================================


def get_tensor_shapes(*, seq_length, micro_batch_size, decoder_seq_length, config, tp_group=None, cp_group=None):
    """Shape of the activation crossing a stage boundary: ``[s/(cp·tp if SP), b, h]``."""
    cp = get_pg_size(cp_group) if cp_group is not None else ps.get_context_parallel_world_size()
    tp = get_pg_size(tp_group) if tp_group is not None else ps.get_tensor_model_parallel_world_size()
    s = seq_length // cp
    if config.sequence_parallel:
        s //= tp
    return [(s, micro_batch_size, config.hidden_size)]


def forward_backward_pipelining_without_interleaving(*, forward_step_func, data_iterator, model, num_microbatches: int, seq_length: int,
                                                     micro_batch_size: int, decoder_seq_length: Optional[int] = None, forward_only: bool = False,
                                                     collect_non_loss_data: bool = False, first_val_step: Optional[bool] = None,
                                                     adjust_tensor_shapes_fn=None, p2p_communicator: Optional[P2PCommunicator] = None,
                                                     pg_collection=None, force_all_reduce: bool = False):
    """warm-up ``pp - rank - 1`` forwards → steady 1F1B → cool-down backwards."""
    if isinstance(model, list):
        assert len(model) == 1
        model = model[0]
    if isinstance(data_iterator, list):
        data_iterator = data_iterator[0]
    config = get_model_config(model)
    if config.overlap_p2p_comm:
        raise ValueError("non-interleaved pipeline parallelism does not support overlapping p2p communication")
    p2p = p2p_communicator or P2PCommunicator(config=config)
    if config.timers is not None:
        config.timers("forward-backward", log_level=1).start(barrier=config.barrier_with_L1_time)
    no_sync_func = config.no_sync_func or contextlib.nullcontext
    no_sync_context = None

    def disable_grad_sync():
        nonlocal no_sync_context
        if no_sync_context is None:
            no_sync_context = no_sync_func()
            no_sync_context.__enter__()

    def enable_grad_sync():
        nonlocal no_sync_context
        if no_sync_context is not None:
            no_sync_context.__exit__(None, None, None)
            no_sync_context = None

    disable_grad_sync()
    pp, rank = p2p.size, p2p.rank_in_group
    num_warmup = min(pp - rank - 1, num_microbatches)
    num_remaining = num_microbatches - num_warmup
    model_type = get_model_type(model)
    shape = get_tensor_shapes(seq_length=seq_length, micro_batch_size=micro_batch_size, decoder_seq_length=decoder_seq_length, config=config)[0]
    inputs, outputs = [], []
    forward_data_store = []
    total_num_tokens = torch.zeros([], dtype=torch.int)
    is_last = p2p.is_last

    def fwd(i):
        nonlocal total_num_tokens
        x = p2p_in.pop(0)
        out, nt = forward_step(forward_step_func, data_iterator, model, num_microbatches, x, forward_data_store, config,
                               collect_non_loss_data=collect_non_loss_data,
                               is_first_microbatch=check_first_val_step(first_val_step, forward_only, i == 0),
                               current_microbatch=i, is_last_stage=is_last)
        total_num_tokens = total_num_tokens.to(nt.device) + nt
        return x, out

    p2p_in: List = []
    # ---- warm-up -------------------------------------------------------------------------
    for i in range(num_warmup):
        p2p_in.append(p2p.recv_forward(shape))
        x, out = fwd(i)
        p2p.send_forward(out)
        if not forward_only:
            inputs.append(x), outputs.append(out)
            deallocate_output_tensor(out, config.deallocate_pipeline_outputs)
    if num_remaining > 0:
        p2p_in.append(p2p.recv_forward(shape))
    # ---- steady state ----------------------------------------------------------------------
    for i in range(num_remaining):
        last_iter = i == num_remaining - 1
        x, out = fwd(i + num_warmup)
        if forward_only:
            p2p.send_forward(out)
            if not last_iter:
                p2p_in.append(p2p.recv_forward(shape))
            continue
        gout = p2p.send_forward_recv_backward(out, shape)
        inputs.append(x), outputs.append(out)
        deallocate_output_tensor(out, config.deallocate_pipeline_outputs)
        x0, o0 = inputs.pop(0), outputs.pop(0)
        if num_warmup == 0 and last_iter:
            if config.grad_sync_func is None or rank == 0:
                enable_grad_sync()
        gin = backward_step(x0, o0, gout, model_type, config)
        if last_iter:
            p2p.send_backward(gin)
        else:
            p2p_in.append(p2p.send_backward_recv_forward(gin, shape))
    # ---- cool-down -------------------------------------------------------------------------
    if not forward_only:
        for i in range(num_warmup):
            if i == num_warmup - 1:
                if config.grad_sync_func is None or rank == 0:
                    enable_grad_sync()
            x0, o0 = inputs.pop(0), outputs.pop(0)
            gout = p2p.recv_backward(shape)
            gin = backward_step(x0, o0, gout, model_type, config)
            p2p.send_backward(gin)
        if no_sync_context is not None:
            enable_grad_sync()
            if config.grad_sync_func is not None:
                config.grad_sync_func(model.parameters())
    _finalize(config, model, total_num_tokens, forward_only, force_all_reduce, pg_collection)
    if config.timers is not None:
        config.timers("forward-backward").stop()
    return forward_data_store


# ==================================================================================
# interleaved 1F1B
# ==================================================================================


def get_schedule_table(num_microbatches: int, num_model_chunks: int, microbatch_group_size_per_vp_stage: int) -> List[Tuple[int, int]]:
    """Virtual micro-batch order: groups of ``g`` micro-batches run through chunk 0, then the
    same group through chunk 1, …  e.g. m=5, v=2, g=3 →
    (0,0)(1,0)(2,0)(0,1)(1,1)(2,1)(3,0)(4,0)(3,1)(4,1)."""
    table = []
    for lo in range(0, num_microbatches, microbatch_group_size_per_vp_stage):
        hi = min(lo + microbatch_group_size_per_vp_stage, num_microbatches)
        for chunk in range(num_model_chunks):
            table.extend((mb, chunk) for mb in range(lo, hi))
    return table


def get_pp_rank_microbatches(num_microbatches, num_model_chunks, microbatch_group_size_per_vp_stage, forward_only=False,
                             overlap_moe_expert_parallel_comm=False, pp_size=None, pp_rank=None):
    """``(total, all_warmup, num_warmup, num_remaining)`` in *virtual* micro-batches."""
    pp = ps.get_pipeline_model_parallel_world_size() if pp_size is None else pp_size
    rank = ps.get_pipeline_model_parallel_rank() if pp_rank is None else pp_rank
    total = num_microbatches * num_model_chunks
    all_warmup = False
    if forward_only:
        warm = total
    else:
        warm = (pp - rank - 1) * 2 + (num_model_chunks - 1) * microbatch_group_size_per_vp_stage
        if overlap_moe_expert_parallel_comm:
            warm += 1
    if warm >= total:
        warm, all_warmup = total, True
    return total, all_warmup, warm, total - warm


def build_interleaved_plan(num_microbatches: int, vp: int, pp: int, rank: int, group: int, forward_only: bool = False):
    """Pure function: the ordered list of ``("F"|"B", virtual_id, microbatch, chunk)`` this rank
    executes.  Backward chunks run in reverse (``vp-1-chunk``)."""
    table = get_schedule_table(num_microbatches, vp, group)
    total, _, warm, remaining = get_pp_rank_microbatches(num_microbatches, vp, group, forward_only, pp_size=pp, pp_rank=rank)
    plan = [("F", k, *table[k]) for k in range(warm)]
    for k in range(remaining):
        plan.append(("F", warm + k, *table[warm + k]))
        mb, ch = table[k]
        plan.append(("B", k, mb, vp - 1 - ch))
    if not forward_only:
        for k in range(remaining, total):
            mb, ch = table[k]
            plan.append(("B", k, mb, vp - 1 - ch))
    return plan


def forward_backward_pipelining_with_interleaving(*, forward_step_func, data_iterator, model, num_microbatches: int, seq_length: int,
                                                  micro_batch_size: int, decoder_seq_length: Optional[int] = None, forward_only: bool = False,
                                                  collect_non_loss_data: bool = False, first_val_step: Optional[bool] = None,
                                                  adjust_tensor_shapes_fn=None, p2p_communicator: Optional[P2PCommunicator] = None,
                                                  pg_collection=None, force_all_reduce: bool = False):
    """Interleaved (virtual-pipeline) 1F1B.  Every executed step exchanges, in ONE p2p group,
    the activation it just produced / the gradient it just produced for the tensors the *next*
    forward / backward step needs — with ``overlap_p2p_comm`` the waits are deferred until the
    consumer so the transfer hides under the following compute."""
    assert isinstance(model, list), "interleaved pipeline parallelism expected model chunking"
    assert all(isinstance(c, torch.nn.Module) for c in model)
    if not isinstance(data_iterator, list):
        data_iterator = [data_iterator] * len(model)
    config = get_model_config(model[0])
    p2p = p2p_communicator or P2PCommunicator(config=config)
    pp, rank, vp = p2p.size, p2p.rank_in_group, len(model)
    group = config.microbatch_group_size_per_vp_stage or pp
    if num_microbatches % group != 0 and num_microbatches % pp != 0 and group == pp:
        pass  # uneven tail groups are handled by get_schedule_table
    if config.timers is not None:
        config.timers("forward-backward", log_level=1).start(barrier=config.barrier_with_L1_time)
    no_sync_func = config.no_sync_func
    if isinstance(no_sync_func, list):
        fs = no_sync_func

        def no_sync_func():
            st = contextlib.ExitStack()
            for f in fs:
                st.enter_context(f())
            return st

    no_sync_func = no_sync_func or contextlib.nullcontext
    no_sync_context = None

    def disable_grad_sync():
        nonlocal no_sync_context
        if no_sync_context is None:
            no_sync_context = no_sync_func()
            no_sync_context.__enter__()

    def enable_grad_sync():
        nonlocal no_sync_context
        if no_sync_context is not None:
            no_sync_context.__exit__(None, None, None)
            no_sync_context = None

    disable_grad_sync()
    model_type = get_model_type(model[0])
    shape = get_tensor_shapes(seq_length=seq_length, micro_batch_size=micro_batch_size, decoder_seq_length=decoder_seq_length, config=config)[0]
    table = get_schedule_table(num_microbatches, vp, group)
    total, all_warmup, warm, remaining = get_pp_rank_microbatches(num_microbatches, vp, group, forward_only, pp_size=pp, pp_rank=rank)

    in_q = [[] for _ in range(vp)]  # per chunk: stage inputs awaiting forward / saved for backward
    saved_in = [[] for _ in range(vp)]
    saved_out = [[] for _ in range(vp)]
    gout_q = [[] for _ in range(vp)]
    forward_data_store = []
    total_num_tokens = torch.zeros([], dtype=torch.int)
    synced_chunks = set()

    def f_chunk(k):
        return table[k][1]

    def b_chunk(k):
        return vp - 1 - table[k][1]

    def needs_recv_fwd(k):
        """Does forward #k on THIS rank take its input from the previous stage?"""
        if k >= total:
            return False
        return not (rank == 0 and f_chunk(k) == 0)

    def produces_send_fwd(k):
        return not (rank == pp - 1 and f_chunk(k) == vp - 1)

    def needs_recv_bwd(k):
        if k >= total:
            return False
        return not (rank == pp - 1 and b_chunk(k) == vp - 1)

    def produces_send_bwd(k):
        return not (rank == 0 and b_chunk(k) == 0)

    def run_forward(k):
        nonlocal total_num_tokens
        mb, ch = table[k]
        ps.set_virtual_pipeline_model_parallel_rank(ch)
        if config.param_sync_func is not None and mb == 0 and ch + 1 < vp and not forward_only:
            fn = config.param_sync_func[ch + 1] if isinstance(config.param_sync_func, list) else config.param_sync_func
            fn(model[ch + 1].parameters())
        first_stage = rank == 0 and ch == 0
        x = None if first_stage else in_q[ch].pop(0)
        last_stage = rank == pp - 1 and ch == vp - 1
        out, nt = forward_step(forward_step_func, data_iterator[ch], model[ch], num_microbatches, x, forward_data_store, config,
                               collect_non_loss_data=collect_non_loss_data,
                               is_first_microbatch=check_first_val_step(first_val_step, forward_only, mb == 0),
                               current_microbatch=mb, vp_stage=ch, is_last_stage=last_stage)
        total_num_tokens = total_num_tokens.to(nt.device) + nt
        if not forward_only:
            saved_in[ch].append(x), saved_out[ch].append(out)
        return out

    def run_backward(k):
        mb, _ = table[k]
        ch = b_chunk(k)
        ps.set_virtual_pipeline_model_parallel_rank(ch)
        # launch this chunk's DP grad sync with its last micro-batch backward
        is_last_for_chunk = all(not (b_chunk(j) == ch) for j in range(k + 1, total))
        if is_last_for_chunk and config.grad_sync_func is None:
            enable_grad_sync()
            synced_chunks.add(ch)
        x, out = saved_in[ch].pop(0), saved_out[ch].pop(0)
        last_stage = rank == pp - 1 and ch == vp - 1
        gout = None if last_stage else gout_q[ch].pop(0)
        gin = backward_step(x, out, gout, model_type, config)
        if is_last_for_chunk and config.grad_sync_func is not None:
            enable_grad_sync()
            fn = config.grad_sync_func[ch] if isinstance(config.grad_sync_func, list) else config.grad_sync_func
            fn(model[ch].parameters())
            synced_chunks.add(ch)
            disable_grad_sync()
        elif is_last_for_chunk:
            disable_grad_sync()
        return gin

    # Which of the PREVIOUS stage's forwards feeds forward #k here?  For rank>0 it is the
    # previous rank's forward #k; for rank 0 chunk c>0 it is the last rank's forward of
    # (mb, c-1).  Sends/receives between a pair are issued in the same relative order on
    # both sides because both walk `table` in order.
    fwd_k, bwd_k = 0, 0

    def step(do_f: bool, do_b: bool):
        """Execute ≤1 forward and ≤1 backward, then one fused exchange."""
        nonlocal fwd_k, bwd_k
        out = gin = None
        kf = kb = None
        if do_f:
            kf = fwd_k
            out = run_forward(kf)
            fwd_k += 1
        if do_b:
            kb = bwd_k
            gin = run_backward(kb)
            bwd_k += 1
        send_next = out if (kf is not None and produces_send_fwd(kf)) else None
        send_prev = gin if (kb is not None and produces_send_bwd(kb)) else None
        # what do the NEXT forward/backward on this rank need?
        recv_prev = do_f and needs_recv_fwd(fwd_k) and _input_available_after(fwd_k)
        recv_next = (do_b or (do_f and fwd_k == warm and remaining > 0 and bwd_k == 0) or (do_f and all_warmup and fwd_k == total)) \
            and needs_recv_bwd(bwd_k) and not forward_only
        fp, fn, _ = p2p.exchange(send_next=send_next, send_prev=send_prev, recv_prev=recv_prev, recv_next=recv_next, tensor_shape=shape)
        if out is not None and not forward_only:
            deallocate_output_tensor(out, config.deallocate_pipeline_outputs and send_next is not None)
        if recv_prev:
            in_q[f_chunk(fwd_k)].append(fp)
        if recv_next:
            gout_q[b_chunk(bwd_k)].append(fn)

    def _input_available_after(k):
        return True

    # ---- prime: first input ---------------------------------------------------------------
    if needs_recv_fwd(0):
        fp, _, _ = p2p.exchange(recv_prev=True, tensor_shape=shape)
        in_q[f_chunk(0)].append(fp)
    for _ in range(warm):
        step(True, False)
    for _ in range(remaining):
        step(True, True)
    if not forward_only:
        for _ in range(total - bwd_k):
            step(False, True)
        enable_grad_sync()
        if config.grad_sync_func is not None:
            for ch in range(vp):
                if ch not in synced_chunks:
                    fn = config.grad_sync_func[ch] if isinstance(config.grad_sync_func, list) else config.grad_sync_func
                    fn(model[ch].parameters())
    ps.set_virtual_pipeline_model_parallel_rank(0)
    _finalize(config, model, total_num_tokens, forward_only, force_all_reduce, pg_collection)
    if config.timers is not None:
        config.timers("forward-backward").stop()
    return forward_data_store
