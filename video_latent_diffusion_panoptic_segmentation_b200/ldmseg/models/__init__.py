from .unet import UNet, UNetOutput
from .vae import GeneralVAESeg

__all__ = ["UNet", "UNetOutput", "GeneralVAESeg"]
