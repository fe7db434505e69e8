"""B200-native drop-in for the reference's `Autoencoder` (AE_pretrained/reconstruction/models/autoencoder.py:233-470):
the same 6-stage encoder-decoder body as `UNet` (skip connections included) with
`reconstruction_output = Sequential(Conv2d(32 -> 3, 3x3, pad 1), Sigmoid())` instead of the 1x1 segmentation head,
trained with `nn.MSELoss` (src/train.py:431) -- BASELINE.json configs[3].

Same constructor keywords (train.py:351-370), attribute tree and `state_dict` keys as the reference; same seed => same
weights.  The whole forward/backward is the fused node of models/unet.py; only the head differs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Type

import torch
import torch.nn as nn

from .unet import ConvBlock, SpatialDropout2d, UNet, UpBlock  # noqa: F401  (re-exported like the reference module)


class Autoencoder(UNet):
    head_kind = "recon3x3"

    def __init__(self, in_channels: int = 3, out_channels: int = 3, n_stages: int = 6,
                 features_per_stage: List[int] = None, kernel_sizes: List[Tuple[int, int]] = None,
                 strides: List[Tuple[int, int]] = None, n_conv_per_stage: List[int] = None,
                 n_conv_per_stage_decoder: List[int] = None, conv_bias: bool = True,
                 norm_op: Type[nn.Module] = nn.InstanceNorm2d, norm_op_kwargs: Dict = None,
                 dropout_op: Optional[Type[nn.Module]] = None, dropout_op_kwargs: Dict = None,
                 nonlin: Type[nn.Module] = nn.LeakyReLU, nonlin_kwargs: Dict = None,
                 encoder_dropout_rates: List[float] = None, decoder_dropout_rates: List[float] = None):
        super().__init__(in_channels=in_channels, num_classes=out_channels, n_stages=n_stages,
                         features_per_stage=features_per_stage, kernel_sizes=kernel_sizes, strides=strides,
                         n_conv_per_stage=n_conv_per_stage, n_conv_per_stage_decoder=n_conv_per_stage_decoder,
                         conv_bias=conv_bias, norm_op=norm_op, norm_op_kwargs=norm_op_kwargs, dropout_op=dropout_op,
                         dropout_op_kwargs=dropout_op_kwargs, nonlin=nonlin, nonlin_kwargs=nonlin_kwargs,
                         encoder_dropout_rates=encoder_dropout_rates, decoder_dropout_rates=decoder_dropout_rates)
        self.out_channels = out_channels

    def _build_head(self, features: int, out_channels: int):
        # autoencoder.py:374-387
        self.reconstruction_output = nn.Sequential(
            nn.Conv2d(features, out_channels, kernel_size=3, stride=1, padding=1, bias=True), nn.Sigmoid())

    def _head_conv(self) -> nn.Conv2d:
        return self.reconstruction_output[0]

    def get_encoder(self):
        """autoencoder.py:438-445"""
        return self.encoder_stages

    def get_decoder(self):
        """autoencoder.py:447-454"""
        return self.decoder_stages, self.reconstruction_output
