from .unet import UNet, UNetOutput
from .vae import GeneralVAESeg
from .vae_image import GeneralVAEImage

__all__ = ["UNet", "UNetOutput", "GeneralVAESeg", "GeneralVAEImage"]
