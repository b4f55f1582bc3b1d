"""Multi-latent attention with matrix absorption (reference ``experimental_attention_variant/absorbed_mla.py``).

Standard MLA expands the cached latent ``c_kv [s, r]`` to per-head keys/values (``k_nope = c_kv W_UK``, ``v = c_kv W_UV``) before attention.
Absorption moves the two up-projections to the QUERY and OUTPUT side, so attention runs directly on the latent — one shared "key/value"
of width ``r + d_rope`` for all heads (MQA-shaped):

    score[h] = (q_nope[h] W_UK[h]ᵀ) · c_kv + q_pe[h] · k_pe          (q absorbed: [d_nope] → [r])
    out[h]   = (softmax(score[h]) · c_kv) W_UV[h]                     (v absorbed on the way out)

Same math (exactly, up to fp rounding), but K/V memory traffic per token drops from ``n·(d_nope + d_rope + d_v)`` to ``r + d_rope``
(DeepSeek-V3: 40 960 → 576 elements), which is what matters for long-context decoding on an HBM-bound GPU, and it is the formulation
sparse attention (``dsa.py``) needs because the top-k gather then touches the small latent rows only.
"""
from __future__ import annotations

import torch

from ..enums import AttnMaskType
from ..multi_latent_attention import MLASelfAttention


class AbsorbedMLASelfAttention(MLASelfAttention):
    """Drop-in for ``MLASelfAttention`` (same parameters / checkpoints); only the attention core differs."""

    def _up_proj_weights(self):
        c = self.config
        w = self.linear_kv_up_proj.weight.view(self.n_local, c.qk_head_dim + c.v_head_dim, c.kv_lora_rank)
        return w[:, : c.qk_head_dim], w[:, c.qk_head_dim :]            # W_UK [n, d_nope, r], W_UV [n, d_v, r]

    def absorbed_qkv(self, hidden_states, inference_context=None):
        """→ (q_abs [sq, b, n, r + d_rope], kv [sk, b, r + d_rope], W_UV) — the tensors a latent-space attention kernel consumes."""
        c = self.config
        q, kv_latent, k_pe = self.get_query_key_value_tensors(hidden_states, inference_context)
        if self.sequence_parallel:
            from ...tensor_parallel.mappings import gather_from_sequence_parallel_region

            kv_latent = gather_from_sequence_parallel_region(kv_latent, group=self.tp_group)
        w_uk, w_uv = self._up_proj_weights()
        q_nope, q_pe = torch.split(q, [c.qk_head_dim, c.qk_pos_emb_head_dim], dim=-1)
        q_lat = torch.einsum("sbnd,ndr->sbnr", q_nope, w_uk.to(q_nope.dtype))
        return torch.cat([q_lat, q_pe], dim=-1), torch.cat([kv_latent, k_pe.squeeze(2)], dim=-1), w_uv

    def latent_attention(self, q_abs, kv, causal: bool, q_offset: int = 0, extra_mask=None):
        """softmax(q_abs · kvᵀ · scale) · kv[..., :r]  → [sq, b, n, r].  ``extra_mask`` [b, sq, sk] bool (True = masked) is OR-ed with the causal mask."""
        r = self.config.kv_lora_rank
        scores = torch.einsum("sbnc,tbc->bnst", q_abs.float(), kv.float()) * self.softmax_scale
        sq, sk = scores.shape[-2:]
        if causal:
            qpos = torch.arange(sq, device=scores.device)[:, None] + q_offset
            scores = scores.masked_fill((torch.arange(sk, device=scores.device)[None, :] > qpos)[None, None], float("-inf"))
        if extra_mask is not None:
            scores = scores.masked_fill(extra_mask[:, None], float("-inf"))
        probs = torch.softmax(scores, dim=-1)
        return torch.einsum("bnst,tbr->sbnr", probs, kv[..., :r].float()).to(q_abs.dtype)

    def forward(self, hidden_states, attention_mask, key_value_states=None, inference_context=None, rotary_pos_emb=None, rotary_pos_cos=None,
                rotary_pos_sin=None, attention_bias=None, packed_seq_params=None, sequence_len_offset=None, *, inference_params=None):
        inference_context = inference_context or inference_params
        c = self.config
        q_abs, kv, w_uv = self.absorbed_qkv(hidden_states, inference_context)
        kv, q_off = self._latent_cache(kv, inference_context)
        causal = self.attn_mask_type == AttnMaskType.causal
        out_lat = self.core_latent_attention(q_abs, kv, causal, q_off, hidden_states)
        out = torch.einsum("sbnr,ndr->sbnd", out_lat, w_uv.to(out_lat.dtype))
        s, b = out.shape[:2]
        return self.linear_proj(out.reshape(s, b, self.n_local * c.v_head_dim))

    def _latent_cache(self, kv, inference_context):
        """Latent KV cache: exactly the [kv_latent | k_pe] rows, nothing is ever expanded per head.  → (all keys so far, offset of the new queries)."""
        if inference_context is None:
            return kv, 0
        kvd = inference_context.key_value_memory_dict
        if self.layer_number not in kvd:
            kvd[self.layer_number] = torch.empty(inference_context.max_sequence_length, inference_context.max_batch_size, kv.shape[-1], dtype=kv.dtype, device=kv.device)
        cache = kvd[self.layer_number]
        s0, b0 = inference_context.sequence_len_offset, inference_context.batch_size_offset
        s1, b1 = s0 + kv.shape[0], b0 + kv.shape[1]
        cache[s0:s1, b0:b1] = kv
        return cache[:s1, b0:b1], s0

    def core_latent_attention(self, q_abs, kv, causal, q_off, hidden_states):
        """Dense latent attention; sparse variants (``DSAMLASelfAttention``) override this."""
        return self.latent_attention(q_abs, kv, causal, q_off)


class DSAMLASelfAttention(AbsorbedMLASelfAttention):
    """Absorbed MLA whose core is DeepSeek sparse attention: an index branch scores every causal key per query, the top-k survive, and the latent attention runs
    on those rows only (reference ``experimental_attention_variant/dsa.py`` ``DSAttention`` inside the MLA layer built by
    ``experimental_attention_variant_module_specs.get_dsa_module_spec_for_backend``).  With ``dsa_indexer_topk_freq > 1`` only every n-th layer owns an indexer;
    the layers in between reuse its indices (handed over through ``config``-scoped storage, the layers of one model run in order)."""

    def __init__(self, config, submodules, layer_number: int = 1, **kw):
        super().__init__(config, submodules, layer_number=layer_number, **kw)
        from .dsa import DSAttention

        self.dsa = DSAttention(config, layer_number=self.layer_number, softmax_scale=self.softmax_scale)

    def core_latent_attention(self, q_abs, kv, causal, q_off, hidden_states):
        from .dsa import source_dsa_compute_layer

        if self.sequence_parallel:
            from ...tensor_parallel.mappings import gather_from_sequence_parallel_region

            hidden_states = gather_from_sequence_parallel_region(hidden_states, group=self.tp_group)
        if kv.shape[0] != hidden_states.shape[0]:
            raise NotImplementedError("DSA with a KV cache needs the index keys cached as well: use the dense absorbed MLA for incremental decoding")
        store = self.config.__dict__.setdefault("_dsa_shared_indices", {})
        shared = None
        if self.dsa.indexer is None:
            shared = store[source_dsa_compute_layer(self.layer_number, self.dsa.skip_offset, self.dsa.topk_freq)]
        out = self.dsa(q_abs, kv, hidden_states, v_width=self.config.kv_lora_rank, q_offset=q_off, shared_indices=shared)
        if self.dsa.indexer is not None:
            store[self.layer_number] = self.dsa.last_indices
        return out
