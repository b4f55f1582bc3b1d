"""Reasoning-block parsers (reference ``parsers/deepseek_r1_reasoning_parser.py``, ``nemotron_v3_reasoning_parser.py``).

``parse(text) -> (content, {"reasoning": …})``: text before the first ``<think>`` is dropped; without a closing ``</think>`` the model is still thinking and
everything is reasoning — unless an implicit end marker (e.g. ``<tool_call>``) shows up first, which ends the reasoning and is kept for the tool parser."""
from __future__ import annotations

from typing import Dict, Tuple


class BaseParser:
    @staticmethod
    def parse(text: str, **kwargs) -> Tuple[str, Dict[str, str]]:
        raise NotImplementedError


class DeepSeekR1ReasoningParser(BaseParser):
    OPEN, CLOSE = "<think>", "</think>"

    @classmethod
    def parse(cls, text: str, **kwargs) -> Tuple[str, Dict[str, str]]:
        before, opened, after = text.partition(cls.OPEN)
        rest = after if opened else before
        end = rest.find(cls.CLOSE)
        start_content = end + len(cls.CLOSE)
        for marker in kwargs.get("implicit_reasoning_end_markers", ()):
            k = rest.find(marker)
            if k >= 0 and (end < 0 or k < end):
                end, start_content = k, k
        if end < 0:
            reasoning, content = rest, ""
        else:
            reasoning, content = rest[:end], rest[start_content:]
        return content, ({"reasoning": reasoning} if reasoning else {})


class NemotronV3ReasoningParser(DeepSeekR1ReasoningParser):
    """Same tags; when thinking is disabled for the request the template emits an EMPTY think block and the whole text is content; a missing ``<think>`` with
    a ``</think>`` present means the opening tag was part of the prompt."""

    @classmethod
    def parse(cls, text: str, **kwargs) -> Tuple[str, Dict[str, str]]:
        if not kwargs.get("enable_thinking", True):
            return text.replace(cls.OPEN, "").replace(cls.CLOSE, "").lstrip("\n"), {}
        if cls.OPEN not in text and cls.CLOSE in text:
            text = cls.OPEN + text
        content, info = super().parse(text, **kwargs)
        return content.lstrip("\n"), ({"reasoning": info["reasoning"].strip("\n")} if info.get("reasoning", "").strip() else {})
