"""Output parsers for generated text (reference ``text/parsers``): reasoning blocks and tool calls."""
from .reasoning import DeepSeekR1ReasoningParser, NemotronV3ReasoningParser  # noqa: F401
from .tool_calls import Qwen3CoderToolParser  # noqa: F401

PARSERS = {"deepseek-r1-reasoning": DeepSeekR1ReasoningParser, "nemotron-v3-reasoning": NemotronV3ReasoningParser, "qwen3-coder-tool": Qwen3CoderToolParser}
