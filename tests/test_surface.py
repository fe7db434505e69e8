"""CPU: the nn.Module surface mirrors the reference's: constructor keywords of train.py:776-795,
attribute tree, state_dict keys/shapes, checkpoint round trip, dropout module placement."""
import io

import torch
import torch.nn as nn

from conftest import load_golden
from unet_implementations_b200.models.losses import SimpleLoss
from unet_implementations_b200.models.unet import ConvBlock, SpatialDropout2d, UNet, UpBlock


def trainer_model():
    # keyword set of Our_UNet/src/train.py:776-795
    return UNet(in_channels=3, num_classes=3, n_stages=6, features_per_stage=[32, 64, 128, 256, 512, 512],
                kernel_sizes=[[3, 3]] * 6, strides=[[1, 1]] + [[2, 2]] * 5, n_conv_per_stage=[2] * 6,
                n_conv_per_stage_decoder=[2] * 5, conv_bias=True, norm_op=nn.InstanceNorm2d,
                norm_op_kwargs={"eps": 1e-5, "affine": True}, dropout_op=None, nonlin=nn.LeakyReLU,
                nonlin_kwargs={"inplace": True}, encoder_dropout_rates=[0.0, 0.0, 0.1, 0.2, 0.3, 0.3],
                decoder_dropout_rates=[0.3, 0.2, 0.2, 0.1, 0.0])


def test_trainer_ctor_and_attribute_tree():
    m = trainer_model()
    assert (m.in_channels, m.num_classes, m.n_stages) == (3, 3, 6)
    assert m.features_per_stage == [32, 64, 128, 256, 512, 512]
    assert isinstance(m.encoder_stages, nn.ModuleList) and len(m.encoder_stages) == 6
    assert isinstance(m.decoder_stages[0], UpBlock) and isinstance(m.decoder_stages[0].conv_block, ConvBlock)
    # Grad-CAM hook target of the reference (utils/visualize.py:457)
    assert isinstance(m.decoder_stages[0].conv_block.block[0], nn.Conv2d)
    assert isinstance(m.encoder_stages[2].block[3], SpatialDropout2d) and m.encoder_stages[2].block[3].drop_prob == 0.1
    assert isinstance(m.encoder_stages[0].block[3], nn.Conv2d)  # no dropout module when rate == 0
    assert tuple(m.segmentation_output.weight.shape) == (3, 32, 1, 1)
    g = load_golden("default_unet_64.pt")
    assert list(m.state_dict().keys()) == g["keys"]
    assert all(v.dtype == torch.float32 for v in m.state_dict().values())


def test_checkpoint_round_trip_and_reference_checkpoint_loads():
    g = load_golden("small_unet.pt")
    m = UNet(**g["cfg"])
    m.load_state_dict(g["state_dict"])  # a state_dict saved by the reference loads strictly
    buf = io.BytesIO()
    torch.save({"model_state_dict": m.state_dict()}, buf)
    buf.seek(0)
    m2 = UNet(**g["cfg"])
    m2.load_state_dict(torch.load(buf)["model_state_dict"])
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


def test_initialize_weights_statistics():
    torch.manual_seed(0)
    m = UNet()
    w = m.encoder_stages[0].block[0].weight
    assert abs(w.std().item() - (2.0 / (32 * 9)) ** 0.5) < 0.01  # kaiming_normal_, fan_out, gain sqrt(2)
    assert all(float(mod.bias.abs().max()) == 0 for mod in m.modules() if isinstance(mod, nn.Conv2d))
    assert all(float((mod.weight - 1).abs().max()) == 0 for mod in m.modules() if isinstance(mod, nn.InstanceNorm2d))


def test_simple_loss_surface():
    fn = SimpleLoss(weight_dice=1.0, weight_ce=1.0, ignore_index=255, smooth=1e-5, class_weights=None,
                    dynamic_weights=True)
    assert isinstance(fn, nn.Module) and isinstance(fn.ce, nn.CrossEntropyLoss)
    assert (fn.weight_dice, fn.weight_ce, fn.ignore_index, fn.smooth, fn.dynamic_weights) == (1.0, 1.0, 255, 1e-5, True)


def test_spatial_dropout_reference_semantics_on_cpu():
    # the module itself is plain torch (it is also the mask source of the fused path): eval = identity,
    # train = per-(n,c) Bernoulli keep mask scaled by 1/(1-p), same draw as nn.Dropout2d's seed behaviour
    d = SpatialDropout2d(0.3)
    x = torch.ones(4, 16, 5, 5)
    d.eval()
    assert d(x) is x
    d.train()
    torch.manual_seed(5)
    y = d(x)
    torch.manual_seed(5)
    ref = x.new_empty(4, 16, 1, 1).bernoulli_(0.7).div_(0.7)
    assert torch.equal(y, x * ref)
    vals = set(y.unique().tolist())
    assert vals <= {0.0, 1.0 / 0.7} or vals <= {0.0, float(torch.tensor(1.0) / 0.7)}
