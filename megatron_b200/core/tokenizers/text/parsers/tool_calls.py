"""Tool-call parser for the Qwen3-Coder XML-ish format (reference ``parsers/qwen3_coder_tool_parser.py``):

    <tool_call>
    <function=NAME>
    <parameter=KEY>
    VALUE
    </parameter>
    …
    </function>
    </tool_call>

``parse(text, tools=[…]) -> (content without the calls, {"tool_calls": [{"name", "arguments": {…}}, …]})``.  Parameter values are converted with the
JSON-schema type from the tool definition when one is given (integer / number / boolean / object / array), else kept as strings; an unterminated last
call (generation cut off) is parsed as far as it goes."""
from __future__ import annotations

import json
import re
from typing import Any, Dict, List, Optional, Tuple

from .reasoning import BaseParser

_CALL = re.compile(r"<tool_call>(.*?)(?:</tool_call>|$)", re.S)
_FUNC = re.compile(r"<function=([^>\n]+)>(.*?)(?:</function>|$)", re.S)
_PARAM = re.compile(r"<parameter=([^>\n]+)>(.*?)(?:</parameter>|(?=<parameter=)|$)", re.S)


def _schema_type(tools: Optional[List[dict]], fn: str, key: str) -> Optional[str]:
    for t in tools or []:
        f = t.get("function", t)
        if f.get("name") == fn:
            return ((f.get("parameters") or {}).get("properties") or {}).get(key, {}).get("type")
    return None


def _convert(value: str, typ: Optional[str]) -> Any:
    v = value.strip("\n")
    if typ in ("integer", "int"):
        try:
            return int(v.strip())
        except ValueError:
            return v
    if typ in ("number", "float"):
        try:
            return float(v.strip())
        except ValueError:
            return v
    if typ in ("boolean", "bool"):
        return v.strip().lower() == "true"
    if typ in ("object", "array") or (typ is None and v.strip()[:1] in "[{"):
        try:
            return json.loads(v)
        except json.JSONDecodeError:
            return v
    if typ is None and v.strip().lower() in ("null", "none"):
        return None
    return v


class Qwen3CoderToolParser(BaseParser):
    @staticmethod
    def parse(text: str, **kwargs) -> Tuple[str, Dict[str, Any]]:
        tools = kwargs.get("tools")
        calls = []
        for block in _CALL.findall(text):
            for name, body in _FUNC.findall(block):
                name = name.strip()
                args = {k.strip(): _convert(v, _schema_type(tools, name, k.strip())) for k, v in _PARAM.findall(body)}
                calls.append({"name": name, "arguments": args})
        content = _CALL.sub("", text).strip()
        return content, ({"tool_calls": calls} if calls else {})
