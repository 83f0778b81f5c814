"""A small reader for the gin syntax the reference's configs use (configs/h_rqvae_*.gin), because gin-config is
not available here and the trainer must stay configurable by the same files.

Supported (everything the HiD-VAE gin files contain):
    import a.b.c                                  -> importlib.import_module (errors are reported, not fatal)
    scope.param = <python literal>                 ints, floats, strings, booleans, None, lists, tuples, dicts
    scope.param = %module.path.Enum.MEMBER         constants registered with @constants_from_enum
    # comments, blank lines, whitespace around '='
`@configurable` functions receive the bound values as keyword defaults (explicit call arguments win), like
gin.configurable.  `parse_config_file` mirrors gin.parse_config_file (reference modules/utils.py:58-62)."""
from __future__ import annotations

import ast
import functools
import importlib
import inspect
import re
from typing import Any, Callable, Dict, Optional

_BINDINGS: Dict[str, Dict[str, Any]] = {}
_CONSTANTS: Dict[str, Any] = {}
_CONFIGURABLES: Dict[str, Callable] = {}
_LINE = re.compile(r"^([A-Za-z_][\w./]*)\.([A-Za-z_]\w*)\s*=\s*(.+)$")


class GinLiteError(ValueError):
    pass


def constants_from_enum(cls=None, *, module: Optional[str] = None):
    """Register every member of an Enum as `%<module>.<Enum>.<MEMBER>` (gin.constants_from_enum)."""
    def register(enum_cls):
        mod = module or enum_cls.__module__
        for member in enum_cls:
            _CONSTANTS[f"{mod}.{enum_cls.__name__}.{member.name}"] = member
            _CONSTANTS[f"{enum_cls.__name__}.{member.name}"] = member
        return enum_cls
    return register if cls is None else register(cls)


def constant(name: str, value: Any) -> None:
    _CONSTANTS[name] = value


def configurable(fn=None, *, name: Optional[str] = None):
    """Bind `<name>.<param> = value` lines to the keyword parameters of `fn` (gin.configurable)."""
    def wrap(f):
        scope = name or f.__name__
        params = inspect.signature(f).parameters
        accepts_kwargs = any(p.kind == p.VAR_KEYWORD for p in params.values())

        @functools.wraps(f)
        def bound(*args, **kwargs):
            supplied = set(kwargs) | set(list(params)[: len(args)])
            for key, value in _BINDINGS.get(scope, {}).items():
                if key in supplied:
                    continue
                if key not in params and not accepts_kwargs:
                    raise GinLiteError(f"gin binding {scope}.{key} does not match any parameter of {f.__qualname__}")
                kwargs[key] = value
            return f(*args, **kwargs)

        _CONFIGURABLES[scope] = bound
        bound.__gin_scope__ = scope
        return bound
    return wrap if fn is None else wrap(fn)


def _strip_comment(line: str) -> str:
    out, quote = [], None
    for ch in line:
        if quote:
            out.append(ch)
            if ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
            out.append(ch)
        elif ch == "#":
            break
        else:
            out.append(ch)
    return "".join(out).strip()


def _value(text: str, where: str) -> Any:
    text = text.strip()
    if text.startswith("%"):
        key = text[1:]
        if key not in _CONSTANTS:
            raise GinLiteError(f"{where}: unknown constant %{key} (known: {sorted(_CONSTANTS)[:8]} ...)")
        return _CONSTANTS[key]
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError) as e:
        raise GinLiteError(f"{where}: cannot parse value {text!r}: {e}") from None


def parse_config(text: str, source: str = "<string>", skip_unknown_imports: bool = True) -> None:
    pending = ""
    for lineno, raw in enumerate(text.splitlines(), 1):
        line = _strip_comment(raw)
        if not line and not pending:
            continue
        line = (pending + " " + line).strip() if pending else line
        # a value may span lines while brackets are open
        if sum(line.count(c) for c in "([{") > sum(line.count(c) for c in ")]}"):
            pending = line
            continue
        pending = ""
        where = f"{source}:{lineno}"
        if line.startswith("import ") or line.startswith("from "):
            mod = line.split()[1]
            try:
                importlib.import_module(mod)
            except ImportError as e:
                if not skip_unknown_imports:
                    raise GinLiteError(f"{where}: cannot import {mod}: {e}") from None
            continue
        m = _LINE.match(line)
        if not m:
            raise GinLiteError(f"{where}: cannot parse line {line!r}")
        scope, key, value = m.group(1), m.group(2), m.group(3)
        _BINDINGS.setdefault(scope.split("/")[-1], {})[key] = _value(value, where)
    if pending:
        raise GinLiteError(f"{source}: unbalanced brackets at end of file")


def parse_config_file(path: str, skip_unknown_imports: bool = True) -> None:
    with open(path) as f:
        parse_config(f.read(), source=path, skip_unknown_imports=skip_unknown_imports)


def bind_parameter(name: str, value: Any) -> None:
    scope, key = name.rsplit(".", 1)
    _BINDINGS.setdefault(scope, {})[key] = value


def query_parameter(name: str) -> Any:
    scope, key = name.rsplit(".", 1)
    return _BINDINGS[scope][key]


def clear_config() -> None:
    _BINDINGS.clear()


def bindings(scope: str) -> Dict[str, Any]:
    return dict(_BINDINGS.get(scope, {}))
