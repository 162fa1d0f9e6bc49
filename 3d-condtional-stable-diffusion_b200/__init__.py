"""b200dm -- B200-native (sm_100a) implementation of the reference's sampling hot path:
reverse DDPM sampling of the (conditional) 3D latent U-Net + VQ quantize + 3D decoder.

The directory name follows the reference repo and is not a Python identifier; import it through
the top-level shim:  ``import b200dm``.
"""
from . import _lib, ops, weights  # noqa: F401
from .unet import build_model, UNet, param_spec  # noqa: F401
from .diffusion import DiffusionModel, ConditionalDiffusionModel, Betas  # noqa: F401
from .first_stage import VQVAE, VQGAN, VectorQuantizer, MonaiDecoder, MonaiEncoder, AttnCpDecoder, VqganFamilyDecoder  # noqa: F401
