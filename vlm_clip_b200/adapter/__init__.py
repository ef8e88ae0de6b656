"""Drop-in mirrors of the reference's `adapter` package (adapter/clip_adapter.py, adapter/peclip.py)."""
from .clip_adapter import SharedMHSAttentionAdapter, TextAdapter, VisionAdapter  # noqa: F401
from .peclip import ContextAdapter, SharedAdapter, TextualAdapter  # noqa: F401
