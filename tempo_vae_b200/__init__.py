"""tempo_vae_b200 — B200-native (sm_100a) engine for the TEMPO-VAE train / encode hot path.

Drop-in names (same signatures as the reference's src/model.py, src/model_with_l2.py, src/train_utils.py,
src/tempo_data*.py):

    from tempo_vae_b200 import get_model, SpectralVAE, AutoencoderKL, DiagonalGaussianDistribution
    from tempo_vae_b200 import VAEWithL2Supervision, L2PredictionHead
    from tempo_vae_b200 import Trainer, L2SupervisedTrainer, seed_all, get_device
    from tempo_vae_b200 import TEMPODataLoader, TEMPODataLoaderWithL2

Importing this package loads libtvae_b200.so and fails loudly if it is missing (no CPU / PyTorch fallback).
"""
from ._lib import EXPORTED, LIB_PATH, TvaeError, lib  # noqa: F401  (loads the shared library)
from .model import (ENGINE, AttnBlock, AutoencoderKL, Conv2d, ConvTranspose2d, Decoder,  # noqa: F401
                    DiagonalGaussianDistribution, Encoder, GroupNorm, ResNetBlock, ResNetDown, ResNetUp, SpectralVAE,
                    get_conv, get_model, get_precision, set_precision, zero_init)
from .inference import (GranuleGraph, encode_granule_whole, encode_patches, evaluate_reconstruction, granule_to_patches,  # noqa: F401
                        normalize_radiance, reconstruct_granule_whole)
from .model_with_l2 import L2PredictionHead, VAEWithL2Supervision  # noqa: F401
from .optim import FusedAdamW  # noqa: F401
from .probes import LinearProbe, MLPProbe, probe_metrics, train_probe  # noqa: F401
from .probe_targets import (component_stats, component_targets, nan_median, normalize_component,  # noqa: F401
                            pool_component, sample_probe_pairs)
from .tempo_data import (DevicePrefetcher, DeviceTileCache, HostTileStore, RandomBuffer, TEMPODataLoader,  # noqa: F401
                         TEMPODataset, load_normalization_stats)
from .tile_prep import SpectrumStats, draw_tile_specs, extract_tiles, granule_statistics, process_granule  # noqa: F401
from .tempo_data_with_l2 import DeviceTileCacheWithL2, TEMPODataLoaderWithL2, TEMPODatasetWithL2  # noqa: F401
from .train_utils import L2SupervisedTrainer, Trainer, get_device, get_sqrt_schedule, seed_all  # noqa: F401

__version__ = "0.1.0"
