from .HookedViTConfig import HookedViTConfig
