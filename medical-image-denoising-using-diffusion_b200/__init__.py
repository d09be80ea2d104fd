"""B200-native (sm_100a) drop-in for the /denoise inference hot path of
KushalChaudhari-16/Medical-Image-Denoising-Using-Diffusion.

The directory name contains hyphens, so import it with
``importlib.import_module("medical-image-denoising-using-diffusion_b200")`` or
through the alias module ``xrd_b200`` at the repository root.
"""
from ._lib import XrdError, LIB_PATH, load as load_library          # noqa: F401
from .models import (                                                # noqa: F401
    AttentionBlock, DiffusionDenoiser, EnhancedNAFNet, ExpertDenoiser, FusionModule, HybridDenoisingRouter,
    LayerNorm, NAFBlock, NoiseAnalyzer, ResidualBlock, SimpleGate, SinusoidalPositionEmbeddings,
    UNetDiffusion, ddim_timestep_indices, native_kernel_launches,
)
from .build import build_library                                    # noqa: F401
from .sharding import shard_bounds, shard_batch, gather_outputs, run_sharded   # noqa: F401
from .tiling import tile_plan, extract_tiles, blend_tiles, denoise_tiled          # noqa: F401
from .serving import MicroBatcher                                               # noqa: F401
from .imageio import resize_bicubic_u8, preprocess_u8, postprocess_u8, resample_table   # noqa: F401

__all__ = [
    "XrdError", "LIB_PATH", "load_library", "build_library",
    "UNetDiffusion", "DiffusionDenoiser", "EnhancedNAFNet", "ExpertDenoiser", "NoiseAnalyzer", "FusionModule",
    "HybridDenoisingRouter", "ddim_timestep_indices", "native_kernel_launches",
    "shard_bounds", "shard_batch", "gather_outputs", "run_sharded",
    "tile_plan", "extract_tiles", "blend_tiles", "denoise_tiled", "MicroBatcher",
    "resize_bicubic_u8", "preprocess_u8", "postprocess_u8", "resample_table",
]
