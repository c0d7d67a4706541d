"""Latent-map spellings accepted by Encoder / Decoder (reference models/const.py:1-2)."""
LOCAL_MIX = ("local", "local_mix", "node", "node_mix")
GLOBAL_MIX = ("global", "global_mix", "graph", "graph_mix")
EPS = 1e-16  # reference utils/const.py:5 (compiled into the kernels as GJ_EPS)
