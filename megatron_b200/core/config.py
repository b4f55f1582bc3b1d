"""Process-wide switch for experimental code paths (reference ``core/config.py``; ``--enable-experimental``)."""
ENABLE_EXPERIMENTAL = False


def set_experimental_flag(flag: bool) -> None:
    global ENABLE_EXPERIMENTAL
    ENABLE_EXPERIMENTAL = bool(flag)


def is_experimental_enabled() -> bool:
    return ENABLE_EXPERIMENTAL
