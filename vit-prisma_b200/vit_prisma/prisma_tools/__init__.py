from .hook_point import HookPoint
from .hooked_root_module import HookedRootModule
from .activation_cache import ActivationCache
from .factored_matrix import FactoredMatrix
