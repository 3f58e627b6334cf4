"""Error types mirroring yagi's `enum Error` (src/error.rs:6-14)."""
from __future__ import annotations


class YagiError(Exception):
    """Base of all errors raised by yagi_b200 (Rust: `crate::error::Error`)."""
    code = -1


class InternalError(YagiError):
    code = 1


class ConfigError(YagiError):
    code = 2


class ValueError_(YagiError, ValueError):
    code = 3


class RangeError(YagiError):
    code = 4


class ModeError(YagiError):
    code = 5


class NoConvergenceError(YagiError):
    code = 6


_BY_CODE = {c.code: c for c in (InternalError, ConfigError, ValueError_, RangeError, ModeError, NoConvergenceError)}


def from_status(code: int, message: str) -> YagiError:
    return _BY_CODE.get(code, InternalError)(message)
