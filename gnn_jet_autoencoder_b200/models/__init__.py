from .graphnet import GraphNet
from .encoder import Encoder
from .decoder import Decoder
from .const import LOCAL_MIX, GLOBAL_MIX

__all__ = ["GraphNet", "Encoder", "Decoder", "LOCAL_MIX", "GLOBAL_MIX"]
