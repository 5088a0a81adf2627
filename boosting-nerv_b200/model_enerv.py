"""Drop-in for the reference's model_enerv.py (train_nerv_all.py:17 imports ENeRV_Boost from here)."""
from bnerv_b200.models import ENeRV_Boost  # noqa: F401
from bnerv_b200.layers import Attention, Conv_Up_Block, FeedForward, PreNorm, TransformerBlock  # noqa: F401
