from ..speculative import MTPDraft, SpeculativeDecoder, verify_draft_tokens, verify_draft_tokens_batched  # noqa: F401
