"""gigagan_pytorch_b200 — B200 (sm_100a) native GigaGAN generator/discriminator training path, drop-in for the
class API of lucidrains/gigagan-pytorch (GigaGAN / Generator / Discriminator / AdaptiveConv2DMod / StyleNetwork)."""
from .modules import (AdaptiveConv2DMod, Attend, CrossAttention, CrossAttentionBlock, TextEncoder, Discriminator, Generator, SelfAttention,  # noqa: F401
                      SelfAttentionBlock, StyleNetwork, UnetUpsampler, compute_dtype, set_compute_dtype)

from .trainer import GigaGAN, get_optimizer  # noqa: F401,E402

__version__ = "0.1.0"
