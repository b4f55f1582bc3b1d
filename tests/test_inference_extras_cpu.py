

def test_batched_speculative_verify_reference_semantics():
    """CPU formulation of the batched verify kernel: identical distributions accept everything and sample the bonus row; a greedy draft that disagrees with a
    one-hot target is cut at the first mismatch and corrected to the target's token; the follow-up token comes from the residual distribution."""
    import torch

    from megatron_b200.core.inference.speculative import verify_draft_tokens_batched

    torch.manual_seed(0)
    B, k, V = 6, 4, 50
    tp = torch.softmax(torch.randn(B, k + 1, V), -1)
    dt = torch.multinomial(tp[:, :k].reshape(-1, V), 1).view(B, k)
    n, nxt = verify_draft_tokens_batched(dt, tp[:, :k].clone(), tp)
    assert torch.equal(n, torch.full((B,), k)) and ((nxt >= 0) & (nxt < V)).all()
    target_tok = tp.argmax(-1)
    onehot = torch.nn.functional.one_hot(target_tok, V).float()
    draft = target_tok[:, :k].clone()
    draft[2, 1] = (draft[2, 1] + 1) % V
    draft[4, 0] = (draft[4, 0] + 1) % V
    n, nxt = verify_draft_tokens_batched(draft, None, onehot)
    assert n.tolist() == [4, 4, 1, 4, 0, 4]
    assert torch.equal(nxt, target_tok[torch.arange(B), n])
    # residual sampling never returns a token whose residual mass is zero
    dp = torch.softmax(torch.randn(B, k, V), -1)
    dt = torch.multinomial(dp.view(-1, V), 1).view(B, k)
    for _ in range(20):
        n, nxt = verify_draft_tokens_batched(dt, dp, tp)
        for b in range(B):
            if n[b] < k:
                assert tp[b, n[b], nxt[b]] > dp[b, n[b], nxt[b]]
