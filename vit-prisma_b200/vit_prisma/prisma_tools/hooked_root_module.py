"""HookedRootModule -- name registry + hook lifetime management for hooked models.

Surface and semantics follow reference
src/vit_prisma/prisma_tools/hooked_root_module.py:22-332:

* ``setup()`` walks ``named_modules`` once, stamps ``module.name`` and fills
  ``mod_dict`` / ``hook_dict`` (registration order == definition order, which is
  what fixes the *key order* of a full cache);
* ``hooks(...)`` is a re-entrant context manager; every nesting level tags the
  hooks it adds with ``context_level`` and, on exit (also on exceptions), strips
  exactly that level's non-permanent hooks;
* ``run_with_hooks`` / ``run_with_cache`` / ``get_caching_hooks`` /
  ``add_caching_hooks`` keep the reference signatures, including
  ``names_filter`` as None | str | list | callable, ``remove_batch_dim``,
  ``device`` and ``incl_bwd``.

The subclass (``HookedViT``) overrides ``run_with_cache`` to divert eligible
calls to the fused CUDA chain; everything here is the general, hook-by-hook
path that arbitrary user hooks need.
"""
from __future__ import annotations

import logging
from contextlib import contextmanager
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch.nn as nn

from vit_prisma.prisma_tools.hook_point import HookPoint

NamesFilter = Optional[Union[Callable[[str], bool], Sequence[str], str]]
HookSpec = Tuple[Union[str, Callable[[str], bool]], Callable]


def normalise_names_filter(names_filter: NamesFilter) -> Callable[[str], bool]:
    """None -> accept all; str -> equality; list -> membership; callable -> as is."""
    if names_filter is None:
        return lambda name: True
    if type(names_filter) == str:
        wanted = names_filter
        return lambda name: name == wanted
    if type(names_filter) == list:
        wanted_list = names_filter
        return lambda name: name in wanted_list
    return names_filter


class HookedRootModule(nn.Module):
    def __init__(self, *args):
        super().__init__()
        self.is_caching = False
        self.context_level = 0

    # ---------------------------------------------------------------- registry
    def setup(self) -> None:
        """Call at the end of the model's ``__init__`` once all layers exist."""
        self.mod_dict: Dict[str, nn.Module] = {}
        self.hook_dict: Dict[str, HookPoint] = {}
        for name, module in self.named_modules():
            if not name:
                continue
            module.name = name
            self.mod_dict[name] = module
            if isinstance(module, HookPoint):
                self.hook_dict[name] = module

    def hook_points(self):
        return self.hook_dict.values()

    def _matching_hook_points(self, selector):
        """Yield (name, HookPoint) for a str name or a predicate over names."""
        if type(selector) == str:
            yield selector, self.mod_dict[selector]
        else:
            for hp_name, hp in self.hook_dict.items():
                if selector(hp_name):
                    yield hp_name, hp

    # ---------------------------------------------------------------- removal
    def remove_all_hook_fns(self, dir="both", including_permanent=False, level=None) -> None:
        for hp in self.hook_points():
            hp.remove_hooks(dir, including_permanent, level)

    def clear_context(self) -> None:
        for hp in self.hook_points():
            hp.clear_context()

    def reset_hooks(self, clear_contexts=True, direction="both",
                    including_permanent=False, level=None) -> None:
        if clear_contexts:
            self.clear_context()
        self.remove_all_hook_fns(direction, including_permanent, level)
        self.is_caching = False

    # --------------------------------------------------------------- addition
    def check_hooks_to_add(self, hook_point, hook_point_name, hook, dir="fwd",
                           is_permanent=False, prepend=False) -> None:
        """Subclass veto point (HookedViT refuses hooks whose cfg toggle is off)."""

    def check_and_add_hook(self, hook_point, hook_point_name, hook, dir="fwd",
                           is_permanent=False, level=None, prepend=False) -> None:
        self.check_hooks_to_add(hook_point, hook_point_name, hook, dir=dir,
                                is_permanent=is_permanent, prepend=prepend)
        hook_point.add_hook(hook, dir=dir, is_permanent=is_permanent, level=level, prepend=prepend)

    def add_hook(self, name, hook, dir="fwd", is_permanent=False, level=None, prepend=False) -> None:
        for hp_name, hp in self._matching_hook_points(name):
            self.check_and_add_hook(hp, hp_name, hook, dir=dir, is_permanent=is_permanent,
                                    level=level, prepend=prepend)

    def add_perma_hook(self, name, hook, dir="fwd") -> None:
        self.add_hook(name, hook, dir, is_permanent=True)

    # ---------------------------------------------------------------- contexts
    @contextmanager
    def hooks(self, fwd_hooks: List[HookSpec] = [], bwd_hooks: List[HookSpec] = [],
              reset_hooks_end: bool = True, clear_contexts: bool = True):
        self.context_level += 1
        try:
            for direction, specs in (("fwd", fwd_hooks), ("bwd", bwd_hooks)):
                for selector, hook in specs:
                    for _, hp in self._matching_hook_points(selector):
                        hp.add_hook(hook, dir=direction, level=self.context_level)
            yield self
        finally:
            if reset_hooks_end:
                self.reset_hooks(clear_contexts=clear_contexts, including_permanent=False,
                                 level=self.context_level)
            self.context_level -= 1

    def run_with_hooks(self, *model_args, fwd_hooks: List[HookSpec] = [],
                       bwd_hooks: List[HookSpec] = [], reset_hooks_end: bool = True,
                       clear_contexts: bool = False):
        if len(bwd_hooks) > 0 and reset_hooks_end:
            logging.warning(
                "WARNING: Hooks will be reset at the end of run_with_hooks. "
                "This removes the backward hooks before a backward pass can occur."
            )
        with self.hooks(fwd_hooks, bwd_hooks, reset_hooks_end, clear_contexts) as hooked:
            return hooked.forward(*model_args)

    # ----------------------------------------------------------------- caching
    def _make_savers(self, cache: dict, device, remove_batch_dim: bool):
        def keep(t):
            t = t.detach().to(device)
            return t[0] if remove_batch_dim else t

        def save_fwd(tensor, hook):
            cache[hook.name] = keep(tensor)

        def save_bwd(tensor, hook):
            cache[hook.name + "_grad"] = keep(tensor)

        return save_fwd, save_bwd

    def get_caching_hooks(self, names_filter: NamesFilter = None, incl_bwd: bool = False,
                          device=None, remove_batch_dim: bool = False,
                          cache: Optional[dict] = None) -> Tuple[dict, list, list]:
        """Return ``(cache, fwd_hooks, bwd_hooks)`` without attaching anything."""
        cache = {} if cache is None else cache
        accept = normalise_names_filter(names_filter)
        self.is_caching = True
        save_fwd, save_bwd = self._make_savers(cache, device, remove_batch_dim)
        fwd, bwd = [], []
        for name in self.hook_dict:
            if accept(name):
                fwd.append((name, save_fwd))
                if incl_bwd:
                    bwd.append((name, save_bwd))
        return cache, fwd, bwd

    def add_caching_hooks(self, names_filter: NamesFilter = None, incl_bwd: bool = False,
                          device=None, remove_batch_dim: bool = False,
                          cache: Optional[dict] = None) -> dict:
        """Attach save-hooks directly (level-less, removed by the next reset_hooks)."""
        cache = {} if cache is None else cache
        accept = normalise_names_filter(names_filter)
        self.is_caching = True
        save_fwd, save_bwd = self._make_savers(cache, device, remove_batch_dim)
        for name, hp in self.hook_dict.items():
            if accept(name):
                hp.add_hook(save_fwd, dir="fwd")
                if incl_bwd:
                    hp.add_hook(save_bwd, dir="bwd")
        return cache

    def run_with_cache(self, *model_args, names_filter: NamesFilter = None, device=None,
                       remove_batch_dim: bool = False, incl_bwd: bool = False,
                       reset_hooks_end: bool = True, clear_contexts: bool = False,
                       fwd_hooks: List[HookSpec] = [], bwd_hooks: List[HookSpec] = [],
                       **model_kwargs):
        cache, cache_fwd, cache_bwd = self.get_caching_hooks(
            names_filter, incl_bwd, device, remove_batch_dim=remove_batch_dim
        )
        with self.hooks(fwd_hooks=fwd_hooks + cache_fwd, bwd_hooks=bwd_hooks + cache_bwd,
                        reset_hooks_end=reset_hooks_end, clear_contexts=clear_contexts):
            model_out = self(*model_args, **model_kwargs)
            if incl_bwd or bwd_hooks:
                model_out.backward()
        return model_out, cache
