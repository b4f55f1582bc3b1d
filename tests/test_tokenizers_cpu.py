"""Tokenizer libraries: tiktoken-format BPE (own implementation), chat templates + SFT masking, reasoning / tool-call parsers, multimodal placeholder expansion."""
import base64
import json

from megatron_b200.core.tokenizers.tokenizer import ByteLevelTokenizer, MegatronTokenizer, build_tokenizer


def _write_vocab(path, merges=(b"he", b"ll", b"llo", b"hello", b" w", b"or", b" wor", b"ld")):
    toks = [bytes([i]) for i in range(256)] + list(merges)
    with open(path, "w") as f:
        for r, t in enumerate(toks):
            f.write(f"{base64.b64encode(t).decode()} {r}\n")
    return len(toks)


def test_tiktoken_bpe_roundtrip_and_merges(tmp_path):
    from megatron_b200.core.tokenizers.text.tiktoken_tokenizer import TikTokenTokenizer, bpe_encode_piece, load_tiktoken_ranks

    p = tmp_path / "v.tiktoken"
    n = _write_vocab(p)
    tok = TikTokenTokenizer(str(p), num_special_tokens=8, special_tokens=["<unk>", "<s>", "</s>", "<pad>"])
    assert tok.vocab_size == n + 8
    ids = tok.tokenize("hello world", bos=True, eos=True)
    assert ids[0] == tok.bos and ids[-1] == tok.eos
    assert tok.detokenize(ids, skip_special_tokens=True) == "hello world"
    ranks = load_tiktoken_ranks(str(p))
    assert bpe_encode_piece(b"hello", ranks) == [ranks[b"hello"]]                      # whole-piece hit
    assert bpe_encode_piece(b"hell", ranks) == [ranks[b"he"], ranks[b"ll"]]            # lowest-rank merges first
    # special tokens inside text map to their reserved ids; unicode survives the byte fallback
    s = "a</s>né 漢"
    assert tok.detokenize(tok.tokenize(s)) == s and tok.eos in tok.tokenize(s)
    # JSON vocab format of the reference + the registry entry points
    j = tmp_path / "v.json"
    json.dump([{"rank": r, "token_bytes": base64.b64encode(t).decode(), "token_str": ""} for t, r in ranks.items()], open(j, "w"))
    t2 = build_tokenizer("TikTokenTokenizer", tokenizer_model=str(j), num_special_tokens=8)
    assert t2.tokenize("hello world") == tok.tokenize("hello world")
    MegatronTokenizer.write_metadata(str(p), "tiktoken", num_special_tokens=8)
    t3 = MegatronTokenizer.from_pretrained(str(p), str(tmp_path / "tokenizer_metadata.json"))
    assert t3.tokenize("hello") == tok.tokenize("hello")


def test_chat_template_and_sft_targets():
    from megatron_b200.core.tokenizers.text import ChatTemplate, SFTTokenizer
    from megatron_b200.core.tokenizers.text.sft_tokenizer import IGNORE_INDEX

    base = ByteLevelTokenizer()
    conv = [{"role": "system", "content": "be brief"}, {"role": "user", "content": "2+2?"}, {"role": "assistant", "content": "4"}, {"role": "user", "content": "3+3?"},
            {"role": "assistant", "content": "6"}]
    text = ChatTemplate("chatml").render(conv, add_generation_prompt=True)
    assert text.count("<|im_start|>") == 6 and text.endswith("<|im_start|>assistant\n")
    sft = SFTTokenizer(base, "chatml")
    ids, targets = sft.tokenize_conversation(conv)
    assert len(ids) == len(targets)
    learned = base.detokenize([t for t in targets if t != IGNORE_INDEX])
    assert "4" in learned and "6" in learned and "2+2" not in learned and "be brief" not in learned
    # the label of position i is token i+1, and only inside assistant turns
    for i, t in enumerate(targets):
        assert t == IGNORE_INDEX or t == ids[i + 1]


def test_reasoning_and_tool_parsers():
    from megatron_b200.core.tokenizers.text.parsers import PARSERS

    r1 = PARSERS["deepseek-r1-reasoning"]
    assert r1.parse("junk<think>step 1</think>answer") == ("answer", {"reasoning": "step 1"})
    assert r1.parse("<think>still going") == ("", {"reasoning": "still going"})
    assert r1.parse("<think>plan<tool_call>x", implicit_reasoning_end_markers=("<tool_call>",)) == ("<tool_call>x", {"reasoning": "plan"})
    nm = PARSERS["nemotron-v3-reasoning"]
    assert nm.parse("thoughts\n</think>\nfinal") == ("final", {"reasoning": "thoughts"})
    assert nm.parse("<think></think>hi", enable_thinking=False) == ("hi", {})
    tp = PARSERS["qwen3-coder-tool"]
    text = "I will look.\n<tool_call>\n<function=search>\n<parameter=query>\nb200 hbm\n</parameter>\n<parameter=top_k>\n3\n</parameter>\n<parameter=opts>\n{\"safe\": true}\n</parameter>\n</function>\n</tool_call>"
    tools = [{"type": "function", "function": {"name": "search", "parameters": {"properties": {"query": {"type": "string"}, "top_k": {"type": "integer"}, "opts": {"type": "object"}}}}}]
    content, info = tp.parse(text, tools=tools)
    assert content == "I will look." and info["tool_calls"] == [{"name": "search", "arguments": {"query": "b200 hbm", "top_k": 3, "opts": {"safe": True}}}]
    # generation cut off inside the call: what is there is still parsed
    cut, info2 = tp.parse("<tool_call>\n<function=f>\n<parameter=a>\n1")
    assert info2["tool_calls"][0]["name"] == "f" and info2["tool_calls"][0]["arguments"]["a"] == "1"


def test_multimodal_tokenizer_expands_image_placeholders():
    from megatron_b200.core.tokenizers.vision import MultimodalTokenizer

    mm = MultimodalTokenizer(ByteLevelTokenizer(), num_image_tokens=4, image_token_id=-200, prompt_format="plain")
    ids = mm.tokenize("a<image>b<image>")
    assert ids.count(-200) == 8 and mm.image_positions(ids) == [(1, 5), (6, 10)]
    assert mm.detokenize(ids) == "ab"
    toks, targets = mm.tokenize_conversation([{"role": "user", "content": "<image>what?"}, {"role": "assistant", "content": "cat"}])
    assert -200 in toks and all(t != -200 for t in targets)
    assert isinstance(MegatronTokenizer.from_pretrained(None, {"library": "null-multimodal"}, vocab_size=100, num_image_tokens=2), MultimodalTokenizer)
