"""B200-native drop-in for the reference's CLIP-conditioned `UNet` (CLIP_UNet/models/unet.py:233-492) -- BASELINE.json
configs[4]: the Our_UNet body plus a fusion layer at the bottleneck,

    x = clip_fusion_conv(cat([x, clip_features], 1)),   clip_fusion_conv = Conv2d(512 + clip_dim -> 512, 1x1) + IN + LReLU

(`unet.py:356-364`, `:441-478`), `forward(x, clip_features=None)`.  `clip_features` ([B, clip_dim, 16, 16] patch tokens)
come from a frozen CLIP ViT-B/16 image encoder (`ClipPatchExtractor`, unet.py:494-620: a third-party model, no
gradient) and are an INPUT of this module: any tensor of that shape works, the encoder itself is out of scope.
Same constructor keywords, attribute tree, `state_dict` keys and same-seed weights as the reference.  The fusion layer
runs inside the fused forward/backward of models/unet.py: the last encoder unit writes its activation into the concat
buffer, the 1x1 conv runs on the tensor-core conv kernels as the centre tap of a zero 3x3 kernel.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Type

import torch
import torch.nn as nn

from .unet import ConvBlock, SpatialDropout2d, UpBlock  # noqa: F401
from .unet import UNet as _BaseUNet


class UNet(_BaseUNet):
    def __init__(self, in_channels: int = 3, num_classes: int = 3, n_stages: int = 6,
                 features_per_stage: List[int] = None, kernel_sizes: List[Tuple[int, int]] = None,
                 strides: List[Tuple[int, int]] = None, n_conv_per_stage: List[int] = None,
                 n_conv_per_stage_decoder: List[int] = None, conv_bias: bool = True,
                 norm_op: Type[nn.Module] = nn.InstanceNorm2d, norm_op_kwargs: Dict = None,
                 dropout_op: Optional[Type[nn.Module]] = None, dropout_op_kwargs: Dict = None,
                 nonlin: Type[nn.Module] = nn.LeakyReLU, nonlin_kwargs: Dict = None,
                 encoder_dropout_rates: List[float] = None, decoder_dropout_rates: List[float] = None,
                 with_clip_features: bool = True, clip_dim: int = 512):
        # read by _build_bottleneck, which UNet.__init__ calls between the encoder and the decoder (the reference's
        # construction order, unet.py:326-364: same seed => same weights)
        object.__setattr__(self, "_clip_cfg", (bool(with_clip_features), int(clip_dim)))
        super().__init__(in_channels=in_channels, num_classes=num_classes, n_stages=n_stages,
                         features_per_stage=features_per_stage, kernel_sizes=kernel_sizes, strides=strides,
                         n_conv_per_stage=n_conv_per_stage, n_conv_per_stage_decoder=n_conv_per_stage_decoder,
                         conv_bias=conv_bias, norm_op=norm_op, norm_op_kwargs=norm_op_kwargs, dropout_op=dropout_op,
                         dropout_op_kwargs=dropout_op_kwargs, nonlin=nonlin, nonlin_kwargs=nonlin_kwargs,
                         encoder_dropout_rates=encoder_dropout_rates, decoder_dropout_rates=decoder_dropout_rates)
        self.with_clip_features = bool(with_clip_features)
        self.clip_dim = int(clip_dim)
        self._extra_features = None

    def _build_bottleneck(self, features, conv_bias, norm_op, norm_op_kwargs, nonlin, nonlin_kwargs):
        with_clip, clip_dim = self._clip_cfg
        if with_clip:  # unet.py:356-364
            self.clip_fusion_conv = nn.Sequential(nn.Conv2d(features + clip_dim, features, kernel_size=1, bias=conv_bias),
                                                  norm_op(features, **norm_op_kwargs), nonlin(**nonlin_kwargs))
            self._fusion_adapted = False

    def _fusion_unit(self):
        if not getattr(self, "with_clip_features", False) or not hasattr(self, "clip_fusion_conv"):
            return None
        conv, norm, act = self.clip_fusion_conv[0], self.clip_fusion_conv[1], self.clip_fusion_conv[2]
        return conv, norm, act, None

    def forward(self, x, clip_features=None):
        if self.with_clip_features and clip_features is not None:
            enc_c = self.features_per_stage[-1]
            expected = enc_c + clip_features.shape[1]
            if not self._fusion_adapted and self.clip_fusion_conv[0].in_channels != expected:
                # the reference re-creates the layer for the channel count it meets on the first call (unet.py:459-475)
                print(f"Adapting fusion layer: {self.clip_fusion_conv[0].in_channels} → {expected}")
                self.clip_fusion_conv = nn.Sequential(nn.Conv2d(expected, enc_c, kernel_size=1, bias=True),
                                                      nn.InstanceNorm2d(enc_c, eps=1e-5, affine=True),
                                                      nn.LeakyReLU(inplace=True)).to(x.device)
                self._fusion_adapted = True
            self._extra_features = clip_features.detach()
        else:
            self._extra_features = None
        try:
            return super().forward(x)
        finally:
            self._extra_features = None
