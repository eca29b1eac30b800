"""Import shims so the UNMODIFIED reference (/root/reference/src) imports in this container.

Used only by make_golden.py (fixture generation).  Nothing here is on any product or test path:
the GPU box has no /root/reference.  Three gaps are bridged:
  * ``fancy_einsum`` (not installed): only ``einsum(eq, *tensors)`` with long axis names is used;
  * ``line_profiler`` (not installed): ``profile`` is imported and never applied;
  * ``open_clip``, ``timm``, ``plotly``, ``matplotlib``, ``kaleido``, ``wandb`` ...: imported at module top level by
    loaders / evals / visualisation code that neither hot path calls -- permissive stub packages.
"""
import importlib.abc
import importlib.machinery
import re
import sys
import types

import torch


def _fancy_einsum(equation: str, *tensors):
    lhs, rhs = equation.split("->")
    names = {}

    def conv(term):
        out = ""
        for tok in term.split():
            if tok == "...":
                out += "..."
                continue
            if tok not in names:
                names[tok] = chr(ord("a") + len(names))
            out += names[tok]
        return out

    terms = [conv(t) for t in re.sub(r"\s+", " ", lhs.replace("\\", " ")).split(",")]
    return torch.einsum(",".join(terms) + "->" + conv(re.sub(r"\s+", " ", rhs.replace("\\", " "))), *tensors)


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Stub(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")

    def __mro_entries__(self, bases):
        return (object,)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("open_clip", "timm", "plotly", "matplotlib", "kaleido", "wandb", "line_profiler", "fancy_einsum")

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _Stub(spec.name)
        if spec.name == "fancy_einsum":
            mod.einsum = _fancy_einsum
        if spec.name == "line_profiler":
            mod.profile = lambda f: f
        return mod

    def exec_module(self, module):
        pass


def install(reference_src="/root/reference/src"):
    import importlib.util
    finder = _StubFinder()
    finder.ROOTS = tuple(r for r in finder.ROOTS if importlib.util.find_spec(r) is None)
    sys.meta_path.insert(0, finder)
    if reference_src not in sys.path:
        sys.path.insert(0, reference_src)
