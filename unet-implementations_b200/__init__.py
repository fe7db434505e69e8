"""B200-native (sm_100a) implementation of the Our_UNet training step of Ulixes-8/UNet-Implementations.

Public surface (mirrors the reference's import surface, Our_UNet/src/train.py:28-29):
    from unet_implementations_b200.models.unet import UNet
    from unet_implementations_b200.models.losses import SimpleLoss
or, as a drop-in for the reference tree, put this directory first on PYTHONPATH and `from models.unet import UNet`.
"""
__version__ = "0.1.0"
