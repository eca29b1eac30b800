"""LensHandle -- bookkeeping record for one hook registered on a HookPoint.

Mirrors reference src/vit_prisma/prisma_tools/lens_handle.py:18-28 (fields
``hook``, ``is_permanent``, ``context_level``) because ``HookPoint.fwd_hooks``
is a public list of these records that tests and notebooks introspect.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from torch.utils.hooks import RemovableHandle


@dataclass
class LensHandle:
    hook: RemovableHandle          # torch's handle; .remove() detaches the hook
    is_permanent: bool = False     # survives reset_hooks() unless including_permanent
    context_level: Optional[int] = None  # nesting depth of the hooks() ctx that added it
