"""Schedule plans: a GPT model chunk expressed as independently schedulable nodes
(reference ``models/common/model_chunk_schedule_plan.py`` — ``TransformerLayerSchedulePlan``, ``TransformerModelChunkSchedulePlan``;
node construction in ``models/gpt/fine_grained_callables.py``).

Per layer five nodes:  ``attn`` (compute: norm, attention, residual, pre-MLP norm, router, local permutation) →
``dispatch`` (communication: expert-parallel token exchange) → ``mlp`` (compute: expert GEMMs, shared experts) →
``combine`` (communication) → ``post`` (compute: un-permutation, shared-expert add, residual).  Dense layers use no-op
communication nodes.  ``run(f_plan, b_plan)`` walks the forward plan of one micro-batch and the backward plan of another
layer by layer, always issuing a communication node of one next to a compute node of the other, which is what hides the
expert-parallel all-to-all (on B200: the NVLink push/pull kernels) behind GEMMs.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ...pipeline_parallel.utils import AbstractSchedulePlan, NoopScheduleNode, ScheduleNode, get_comm_stream, get_comp_stream


def _new_event():
    return torch.cuda.Event() if torch.cuda.is_available() and get_comp_stream() is not None else None


class TransformerLayerSchedulePlan:
    """Nodes of one transformer layer for one micro-batch."""

    def __init__(self, layer, event, extra: dict):
        self.layer = layer
        moe = getattr(layer, "is_moe_layer", False)
        mk = lambda fn, comm, name: ScheduleNode(fn, get_comm_stream if comm else get_comp_stream, event, name=f"L{layer.layer_number}.{name}")  # noqa: E731
        self.attn = mk(self._attn_moe if moe else self._attn_dense, False, "attn")
        self.mlp = mk(self._mlp_moe if moe else self._mlp_dense, False, "mlp")
        if moe:
            self.dispatch = mk(self._dispatch, True, "dispatch")
            self.combine = mk(self._combine, True, "combine")
            self.post = mk(self._post_moe, False, "post")
        else:
            self.dispatch, self.combine, self.post = NoopScheduleNode(), NoopScheduleNode(), NoopScheduleNode()
        self.extra = extra
        self._state = {}

    # ---- dense ----
    def _attn_dense(self, hidden):
        e = self.extra
        h, _ = self.layer._forward_attention(hidden, e.get("attention_mask"), None, None, e.get("rotary_pos_emb"), e.get("attention_bias"), None,
                                             e.get("packed_seq_params"))
        return h

    def _mlp_dense(self, hidden):
        return self.layer._forward_mlp(hidden)

    # ---- MoE ----
    def _attn_moe(self, hidden):
        L = self.layer
        h = self._attn_dense(hidden)
        normed = L.pre_mlp_layernorm(h)
        moe = L.mlp
        probs, routing_map = moe.route(normed)
        x, p = moe.token_dispatcher.dispatch_preprocess(normed, routing_map, probs)
        shared_in = normed if moe.use_shared_expert else None
        return (h, x, p) if shared_in is None else (h, x, p, shared_in)

    def _dispatch(self, h, x, p, *shared):
        x, p = self.layer.mlp.token_dispatcher.token_dispatch(x, p)
        return (h, x, p) + tuple(shared)

    def _mlp_moe(self, h, x, p, *shared):
        moe = self.layer.mlp
        x, tokens_per_expert, p = moe.token_dispatcher.dispatch_postprocess(x, p)
        out, bias = moe.experts(x, tokens_per_expert, p)
        self._state["bias"] = bias
        out = moe.token_dispatcher.combine_preprocess(out)
        if shared:
            return h, out, moe.shared_experts(shared[0])
        return h, out

    def _combine(self, h, out, *shared):
        return (h, self.layer.mlp.token_dispatcher.token_combine(out)) + tuple(shared)

    def _post_moe(self, h, out, *shared):
        L = self.layer
        out = L.mlp.token_dispatcher.combine_postprocess(out)
        if shared:
            out = out + shared[0]
        return L.mlp_bda(L.training, L.config.bias_dropout_fusion)((out, self._state.pop("bias", None)), h, L.hidden_dropout)

    # ---- plain (non-overlapped) execution ----
    def forward(self, hidden):
        x = self.attn.forward(hidden)
        x = self.dispatch.forward(x)
        x = self.mlp.forward(x)
        x = self.combine.forward(x)
        return self.post.forward(x)

    def backward(self, grad):
        g = self.post.backward(grad)
        g = self.combine.backward(g)
        g = self.mlp.backward(g)
        g = self.dispatch.backward(g)
        return self.attn.backward(g)


class TransformerModelChunkSchedulePlan(AbstractSchedulePlan):
    """One micro-batch through one GPT model chunk: ``pre`` (embedding, rotary) → layer plans → ``post`` (final norm, head, loss)."""

    def __init__(self, model, input_ids, position_ids, attention_mask, labels=None, loss_mask=None, packed_seq_params=None, loss_func=None):
        self.model = model
        self.event = _new_event()
        if self.event is not None:
            self.event.record(get_comp_stream())
        self.loss_func = loss_func
        self.batch = dict(input_ids=input_ids, position_ids=position_ids, attention_mask=attention_mask, labels=labels, loss_mask=loss_mask,
                          packed_seq_params=packed_seq_params)
        decoder_input, rotary = None, None
        self.extra = dict(attention_mask=attention_mask, packed_seq_params=packed_seq_params, rotary_pos_emb=None)
        self.pre = ScheduleNode(self._pre, get_comp_stream, self.event, name="pre")
        self.layers: List[TransformerLayerSchedulePlan] = [TransformerLayerSchedulePlan(l, self.event, self.extra) for l in model.decoder.layers]
        self.post = ScheduleNode(self._post, get_comp_stream, self.event, name="post")
        self.loss = None

    def _pre(self):
        b = self.batch
        decoder_input, rotary = self.model._preprocess(b["input_ids"], b["position_ids"], packed_seq_params=b["packed_seq_params"])
        if decoder_input is None:
            decoder_input = self.model.decoder.input_tensor
        self.extra["rotary_pos_emb"] = rotary
        return decoder_input

    def _post(self, hidden):
        m, b = self.model, self.batch
        if m.decoder.final_layernorm is not None:
            hidden = m.decoder.final_layernorm(hidden)
        out = m._postprocess(hidden, b["input_ids"], b["position_ids"], b["labels"], self.extra["rotary_pos_emb"], b["loss_mask"], b["attention_mask"],
                             b["packed_seq_params"], None, None)
        if self.loss_func is not None:
            out = self.loss_func(out)
        return out

    # ---- execution ----
    @classmethod
    def run(cls, f_plan: Optional["TransformerModelChunkSchedulePlan"], b_plan: Optional["TransformerModelChunkSchedulePlan"], grad=None):
        """Forward of ``f_plan`` interleaved with backward of ``b_plan`` (either may be ``None``).  Returns the forward output (loss)."""
        nf = len(f_plan.layers) if f_plan is not None else 0
        nb = len(b_plan.layers) if b_plan is not None else 0
        x = f_plan.pre.forward(()) if f_plan is not None else None
        g = b_plan.post.backward(grad if grad is not None else _ones_like_outputs(b_plan.post)) if b_plan is not None else None
        for i in range(max(nf, nb)):
            fl = f_plan.layers[i] if i < nf else None
            bl = b_plan.layers[nb - 1 - i] if i < nb else None
            # the five phases: a communication node of one micro-batch is always issued next to a compute node of the other
            if bl is not None:
                g = bl.post.backward(g)
                g = bl.combine.backward(g)          # comm
            if fl is not None:
                x = fl.attn.forward(x)              # comp
                x = fl.dispatch.forward(x)          # comm
            if bl is not None:
                g = bl.mlp.backward(g)              # comp
                g = bl.dispatch.backward(g)         # comm
            if fl is not None:
                x = fl.mlp.forward(x)               # comp
                x = fl.combine.forward(x)           # comm
            if bl is not None:
                g = bl.attn.backward(g)             # comp
            if fl is not None:
                x = fl.post.forward(x)
        if b_plan is not None:
            b_plan.pre.backward(g)
        out = None
        if f_plan is not None:
            out = f_plan.post.forward(x)
            f_plan.loss = out
        return out


def _ones_like_outputs(node: ScheduleNode):
    return tuple(torch.ones_like(o) if isinstance(o, torch.Tensor) and o.requires_grad else None for o in node.outputs)
