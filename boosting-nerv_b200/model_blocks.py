"""Drop-in for the reference's model_blocks.py: same public names, backed by bnerv_b200 (see bnerv_b200/layers.py)."""
from bnerv_b200.layers import *  # noqa: F401,F403
from bnerv_b200.layers import (ActivationLayer, Block, ConvNeXt, Conv_Up_Block, CustomConv2d, CustomLinear, DownConv,
                               LayerNorm, NeRV_MLP, NeRVBlock, NormLayer, OutImg, PositionEncoding, ResBlock_SFT,
                               SFTLayer, Sin, UpConv, quant_map)
