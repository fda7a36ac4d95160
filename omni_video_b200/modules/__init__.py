from .attention import attention, flash_attention  # noqa: F401
from .model import WanModel  # noqa: F401
from .unified_model import UnifiedWanWithMixedConditionModel  # noqa: F401
from .visual_context_adapter import VisualContextAdapter  # noqa: F401
from .t5 import T5Encoder, T5EncoderModel, umt5_xxl  # noqa: F401
