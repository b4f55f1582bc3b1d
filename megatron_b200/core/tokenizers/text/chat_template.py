"""Chat templates (reference ``text/libraries/chat_template.py``): a Jinja template rendered over a list of ``{"role", "content"}`` turns, tokenised with
the wrapped tokenizer.  ``tokenize_conversation`` additionally returns which tokens were produced by assistant turns (the SFT loss mask)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

CHATML = ("{% for m in messages %}<|im_start|>{{ m['role'] }}\n{{ m['content'] }}<|im_end|>\n{% endfor %}"
          "{% if add_generation_prompt %}<|im_start|>assistant\n{% endif %}")
LLAMA3 = ("{% for m in messages %}<|start_header_id|>{{ m['role'] }}<|end_header_id|>\n\n{{ m['content'] }}<|eot_id|>{% endfor %}"
          "{% if add_generation_prompt %}<|start_header_id|>assistant<|end_header_id|>\n\n{% endif %}")
PLAIN = "{% for m in messages %}{{ m['role'] }}: {{ m['content'] }}\n{% endfor %}{% if add_generation_prompt %}assistant: {% endif %}"
TEMPLATES = {"chatml": CHATML, "llama3": LLAMA3, "plain": PLAIN}


class ChatTemplate:
    def __init__(self, template: str = "chatml"):
        import jinja2

        self.source = TEMPLATES.get(template, template)
        self._tpl = jinja2.Environment(trim_blocks=False, lstrip_blocks=False).from_string(self.source)

    def render(self, messages: List[Dict[str, str]], add_generation_prompt: bool = False, **kw) -> str:
        return self._tpl.render(messages=messages, add_generation_prompt=add_generation_prompt, **kw)

    def tokenize_conversation(self, tokenizer, messages: List[Dict[str, str]], add_generation_prompt: bool = False,
                              train_on_roles: Tuple[str, ...] = ("assistant",)) -> Tuple[List[int], List[int]]:
        """→ (token ids, loss mask).  The mask is found by rendering growing prefixes of the conversation: the tokens a turn ADDS belong to that turn
        (robust to templates whose per-turn text is not a simple concatenation)."""
        ids: List[int] = []
        mask: List[int] = []
        prev = 0
        for i, m in enumerate(messages):
            cur = tokenizer.tokenize(self.render(messages[: i + 1]))
            new = cur[prev:]
            ids += new
            mask += [1 if m["role"] in train_on_roles else 0] * len(new)
            prev = len(cur)
        if add_generation_prompt:
            cur = tokenizer.tokenize(self.render(messages, add_generation_prompt=True))
            ids += cur[prev:]
            mask += [0] * (len(cur) - prev)
        return ids, mask
