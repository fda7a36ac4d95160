"""Convenience alias used by tests: the module-granular forward (see blocks_api)."""
from .blocks_api import dit_forward_modular as forward_module_granular  # noqa: F401
