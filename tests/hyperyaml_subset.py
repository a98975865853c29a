"""A small subset of HyperPyYAML, enough to load the reference's hparams files in tests (hyperpyyaml itself is not
installed here): `!ref <key>` (whole-value references, `<key>` interpolation inside strings, simple arithmetic), `!new:mod.Class`
with mapping / sequence / no arguments, `!name:`, `!apply:`, `!PLACEHOLDER`, and `overrides`.  Resolution is LAZY: only the
keys a test asks for are instantiated, so `!new:` entries whose modules are not importable here (speechbrain loggers,
schedulers ...) never run.  Test infrastructure only."""
from __future__ import annotations

import importlib
import re

import yaml


class _Ref:
    def __init__(self, expr):
        self.expr = expr


class _Call:
    def __init__(self, kind, target, args):
        self.kind, self.target, self.args = kind, target, args


class _Placeholder:
    pass


class _Loader(yaml.SafeLoader):
    pass


def _args_of(loader, node):
    if isinstance(node, yaml.MappingNode):
        return loader.construct_mapping(node, deep=True)
    if isinstance(node, yaml.SequenceNode):
        return loader.construct_sequence(node, deep=True)
    v = loader.construct_scalar(node)
    return None if v in ("", None) else v


_Loader.add_constructor("!ref", lambda l, n: _Ref(l.construct_scalar(n)))
_Loader.add_constructor("!PLACEHOLDER", lambda l, n: _Placeholder())
_Loader.add_constructor("!copy", lambda l, n: _Ref(l.construct_scalar(n)))
for _kind in ("new", "name", "apply", "module"):
    _Loader.add_multi_constructor(f"!{_kind}:", (lambda k: lambda l, suffix, n: _Call(k, suffix, _args_of(l, n)))(_kind))


def _import(path):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


class Hparams:
    """hp = Hparams(text, overrides={...}); hp["wav2vec2"] instantiates that entry (and what it references), once."""

    def __init__(self, text: str, overrides=None, class_map=None):
        for old, new in (class_map or {}).items():  # "only the class string changes"
            text = re.sub(r"(!new:|!name:)" + re.escape(old) + r"(?![\w.])", r"\1" + new, text)
        self.raw = yaml.load(text, Loader=_Loader)
        self.raw.update(overrides or {})
        self.cache = {}

    def keys(self):
        return self.raw.keys()

    def __getitem__(self, key):
        if key not in self.cache:
            self.cache[key] = self._resolve(self.raw[key])
        return self.cache[key]

    def _ref(self, expr: str):
        expr = expr.strip()
        m = re.fullmatch(r"<([\w.]+)>", expr)
        if m:  # the object itself (this is how `modules:` shares instances)
            return self._lookup(m.group(1))
        parts = re.split(r"(<[\w.]+>)", expr)
        vals = [self._lookup(p[1:-1]) if re.fullmatch(r"<[\w.]+>", p) else p for p in parts]
        out = "".join(str(v) for v in vals)
        if all(isinstance(v, (int, float)) or re.fullmatch(r"[\s\d.+\-*/()]*", v) for v in vals) and re.search(r"[+\-*/]", out):
            return eval(out, {"__builtins__": {}})  # arithmetic on numbers only
        return out

    def _lookup(self, dotted):
        key, *rest = dotted.split(".")
        v = self[key]
        for r in rest:
            v = v[r] if isinstance(v, dict) else getattr(v, r)
        return v

    def _resolve(self, v):
        if isinstance(v, _Ref):
            return self._ref(v.expr)
        if isinstance(v, _Placeholder):
            raise ValueError("!PLACEHOLDER must be overridden")
        if isinstance(v, dict):
            return {k: self._resolve(x) for k, x in v.items()}
        if isinstance(v, list):
            return [self._resolve(x) for x in v]
        if isinstance(v, _Call):
            fn = _import(v.target)
            args = self._resolve(v.args)
            a, kw = ([], args) if isinstance(args, dict) else (args if isinstance(args, list) else ([] if args is None else [args]), {})
            if v.kind == "name":
                import functools
                return functools.partial(fn, *a, **kw)
            return fn(*a, **kw)
        return v
