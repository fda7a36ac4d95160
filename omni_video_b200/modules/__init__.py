from .attention import attention, flash_attention  # noqa: F401
from .model import WanModel  # noqa: F401
