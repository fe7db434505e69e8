"""Generates the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE (CPU, fp32).

    python tests/golden/make_golden.py        # needs /root/reference (build container only)

The reference (Ulixes-8/UNet-Implementations) ships no tests or golden vectors, so these fixtures -- outputs of
the reference's own `models.unet.UNet` and `models.losses.SimpleLoss` on seeded inputs -- are what pins the
oracle (oracle/unet_oracle.py) and the module surface.  The GPU box has no /root/reference; it only reads the
.pt files written here.
"""
import hashlib
import os
import sys

import torch

REF = "/root/reference/Our_UNet"
HERE = os.path.dirname(os.path.abspath(__file__))


def sd_sha256(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def make_autoencoder_fixture():
    """small_ae.pt: the UNMODIFIED reference Autoencoder (AE_pretrained/reconstruction/models/autoencoder.py) + nn.MSELoss
    (src/train.py:431) on seeded inputs -- run in a subprocess-free way by loading the module from its file, because its
    package name `models` collides with Our_UNet's."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_autoencoder", "/root/reference/AE_pretrained/reconstruction/models/autoencoder.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = dict(in_channels=3, out_channels=3, n_stages=3, features_per_stage=[32, 64, 64],
               kernel_sizes=[[3, 3]] * 3, strides=[[1, 1], [2, 2], [2, 2]], n_conv_per_stage=[2] * 3,
               n_conv_per_stage_decoder=[2] * 2, encoder_dropout_rates=[0.0, 0.05, 0.15],
               decoder_dropout_rates=[0.15, 0.0])
    torch.manual_seed(27)
    model = mod.Autoencoder(**cfg)
    g = torch.Generator().manual_seed(28)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.2)
    x = torch.rand(2, 3, 32, 48, generator=g)           # images in [0, 1] as the AE trainer feeds them
    target = torch.rand(2, 3, 32, 48, generator=g)
    loss_fn = torch.nn.MSELoss()
    model.train()
    torch.manual_seed(99)
    out = model(x)
    loss = loss_fn(out, target)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.eval()
    with torch.no_grad():
        out_eval = model(x)
    # default architecture: key list and same-seed init hash
    torch.manual_seed(1234)
    full = mod.Autoencoder()
    torch.save({"cfg": cfg, "state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "x": x, "target": target,
                "dropout_seed": 99, "output_train": out.detach(), "loss": loss.detach(), "grads": grads,
                "output_eval": out_eval, "default_keys": list(full.state_dict().keys()),
                "default_sha256": sd_sha256(full.state_dict())}, os.path.join(HERE, "small_ae.pt"))


def make_clip_unet_fixture():
    """small_clip_unet.pt: the UNMODIFIED reference CLIP-conditioned UNet (CLIP_UNet/models/unet.py) + the reference
    SimpleLoss on seeded inputs; the CLIP patch features are a seeded random tensor (the encoder is a frozen third-party
    model whose output is an input of this module)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_clip_unet", "/root/reference/CLIP_UNet/models/unet.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    spec2 = importlib.util.spec_from_file_location("ref_clip_losses", "/root/reference/CLIP_UNet/models/losses.py")
    lmod = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(lmod)
    cfg = dict(in_channels=3, num_classes=3, n_stages=3, features_per_stage=[32, 64, 64],
               kernel_sizes=[[3, 3]] * 3, strides=[[1, 1], [2, 2], [2, 2]], n_conv_per_stage=[2] * 3,
               n_conv_per_stage_decoder=[2] * 2, encoder_dropout_rates=[0.0, 0.1, 0.3],
               decoder_dropout_rates=[0.3, 0.0], with_clip_features=True, clip_dim=32)
    torch.manual_seed(37)
    model = mod.UNet(**cfg)
    g = torch.Generator().manual_seed(38)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.2)
    x = torch.randn(2, 3, 32, 48, generator=g)
    clip = torch.randn(2, 32, 8, 12, generator=g)
    clip_other = torch.randn(2, 32, 5, 7, generator=g)   # a patch grid that has to be resized (unet.py:444-451)
    target = torch.randint(0, 3, (2, 32, 48), generator=g)
    target[torch.rand(2, 32, 48, generator=g) < 0.1] = 255
    loss_fn = lmod.SimpleLoss(weight_dice=1.0, weight_ce=1.0, ignore_index=255, smooth=1e-5, class_weights=None,
                              dynamic_weights=True)
    model.train()
    torch.manual_seed(99)
    logits = model(x, clip)
    loss = loss_fn(logits, target)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.eval()
    with torch.no_grad():
        logits_eval = model(x, clip)
        logits_eval_resized = model(x, clip_other)
        logits_eval_noclip = model(x)
    torch.manual_seed(1234)
    full = mod.UNet()
    torch.save({"cfg": cfg, "state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "x": x, "clip": clip,
                "clip_other": clip_other, "target": target, "dropout_seed": 99, "logits_train": logits.detach(),
                "loss": loss.detach(), "grads": grads, "logits_eval": logits_eval,
                "logits_eval_resized": logits_eval_resized, "logits_eval_noclip": logits_eval_noclip,
                "default_keys": list(full.state_dict().keys()), "default_sha256": sd_sha256(full.state_dict())},
               os.path.join(HERE, "small_clip_unet.pt"))


def main():
    sys.path.insert(0, REF)
    from models.losses import SimpleLoss  # noqa: E402  (the reference's own modules)
    from models.unet import UNet  # noqa: E402

    torch.set_num_threads(8)
    torch.use_deterministic_algorithms(False)

    # ---- 1. tiny UNet: full state_dict + outputs + gradients, train and eval mode
    cfg = dict(in_channels=3, num_classes=3, n_stages=3, features_per_stage=[8, 16, 16],
               kernel_sizes=[[3, 3]] * 3, strides=[[1, 1], [2, 2], [2, 2]], n_conv_per_stage=[2] * 3,
               n_conv_per_stage_decoder=[2] * 2, encoder_dropout_rates=[0.0, 0.1, 0.3],
               decoder_dropout_rates=[0.3, 0.0])
    torch.manual_seed(7)
    model = UNet(**cfg)
    # make every parameter non-trivial (the init sets biases to 0 and norm weights to 1)
    g = torch.Generator().manual_seed(8)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.2)
    x = torch.randn(2, 3, 32, 48, generator=g)
    target = torch.randint(0, 3, (2, 32, 48), generator=g)
    target[torch.rand(2, 32, 48, generator=g) < 0.1] = 255
    loss_fn = SimpleLoss(weight_dice=1.0, weight_ce=1.0, ignore_index=255, smooth=1e-5, class_weights=None,
                         dynamic_weights=True)
    model.train()
    torch.manual_seed(99)
    logits_train = model(x)
    loss = loss_fn(logits_train, target)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.eval()
    with torch.no_grad():
        logits_eval = model(x)
    torch.save({"cfg": cfg, "state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "x": x,
                "target": target, "dropout_seed": 99, "logits_train": logits_train.detach(), "loss": loss.detach(),
                "grads": grads, "logits_eval": logits_eval}, os.path.join(HERE, "tiny_unet.pt"))

    # ---- 1b. small UNet whose channel counts are inside the tensor-core envelope (multiples of 32, stem 3->32)
    cfg = dict(in_channels=3, num_classes=3, n_stages=3, features_per_stage=[32, 64, 64],
               kernel_sizes=[[3, 3]] * 3, strides=[[1, 1], [2, 2], [2, 2]], n_conv_per_stage=[2] * 3,
               n_conv_per_stage_decoder=[2] * 2, encoder_dropout_rates=[0.0, 0.1, 0.3],
               decoder_dropout_rates=[0.3, 0.0])
    torch.manual_seed(17)
    model = UNet(**cfg)
    g = torch.Generator().manual_seed(18)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * 0.2)
    x = torch.randn(2, 3, 32, 48, generator=g)
    target = torch.randint(0, 3, (2, 32, 48), generator=g)
    target[torch.rand(2, 32, 48, generator=g) < 0.1] = 255
    model.train()
    torch.manual_seed(99)
    logits_train = model(x)
    loss = loss_fn(logits_train, target)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    # the reference's own bf16 behaviour (torch.autocast on the CPU): the yardstick for what 16-bit storage costs
    model.zero_grad()
    torch.manual_seed(99)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        logits_bf16 = model(x)
        loss_bf16 = loss_fn(logits_bf16, target)
    loss_bf16.backward()
    grads_bf16 = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.eval()
    with torch.no_grad():
        logits_eval = model(x)
    torch.save({"cfg": cfg, "state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "x": x,
                "target": target, "dropout_seed": 99, "logits_train": logits_train.detach(), "loss": loss.detach(),
                "grads": grads, "logits_eval": logits_eval, "logits_train_bf16": logits_bf16.detach().float(),
                "loss_bf16": loss_bf16.detach().float(),
                "grads_bf16": {k: v.bfloat16() for k, v in grads_bf16.items()}}, os.path.join(HERE, "small_unet.pt"))

    # ---- 2. default UNet(): init hash, small-input logits/loss, per-parameter gradient norms
    torch.manual_seed(1234)
    model = UNet()
    sha = sd_sha256(model.state_dict())
    keys = list(model.state_dict().keys())
    shapes = [tuple(v.shape) for v in model.state_dict().values()]
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 64, 64, generator=g)
    target = torch.randint(0, 3, (2, 64, 64), generator=g)
    target[torch.rand(2, 64, 64, generator=g) < 0.1] = 255
    model.train()
    torch.manual_seed(99)
    logits = model(x)
    loss = loss_fn(logits, target)
    loss.backward()
    gnorm = {k: p.grad.norm().item() for k, p in model.named_parameters()}
    gsample = {k: p.grad.flatten()[:64].clone() for k, p in model.named_parameters()}
    model.eval()
    with torch.no_grad():
        logits_eval = model(x)
    torch.save({"sha256": sha, "keys": keys, "shapes": shapes, "n_params": sum(p.numel() for p in model.parameters()),
                "x_seed": 0, "logits_train": logits.detach(), "logits_eval": logits_eval, "loss": loss.detach(),
                "grad_norms": gnorm, "grad_samples": gsample, "torch": torch.__version__},
               os.path.join(HERE, "default_unet_64.pt"))

    # ---- 3. SimpleLoss cases: value + gradient wrt logits
    cases = {}
    g = torch.Generator().manual_seed(3)
    for name in ["uniform", "absent_class", "one_image_ignored", "static_weights", "no_weights", "dice_heavy"]:
        lg = (torch.randn(3, 3, 24, 40, generator=g) * 2).requires_grad_(True)
        tg = torch.randint(0, 3, (3, 24, 40), generator=g)
        tg[torch.rand(3, 24, 40, generator=g) < 0.15] = 255
        kw = dict(weight_dice=1.0, weight_ce=1.0, ignore_index=255, smooth=1e-5, class_weights=None, dynamic_weights=True)
        if name == "absent_class":
            tg[tg == 2] = 1
        elif name == "one_image_ignored":
            tg[1] = 255
        elif name == "static_weights":
            kw.update(class_weights=torch.tensor([0.5, 1.25, 2.0]), dynamic_weights=False)
        elif name == "no_weights":
            kw.update(dynamic_weights=False)
        elif name == "dice_heavy":
            kw.update(weight_dice=2.5, weight_ce=0.25)
        fn = SimpleLoss(**kw)
        val = fn(lg, tg)
        val.backward()
        kw_s = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items()}
        cases[name] = {"logits": lg.detach().clone(), "target": tg, "kwargs": kw_s, "loss": val.detach(),
                       "dlogits": lg.grad.clone()}
    torch.save(cases, os.path.join(HERE, "loss_cases.pt"))
    make_autoencoder_fixture()
    make_clip_unet_fixture()
    for f in ["tiny_unet.pt", "small_unet.pt", "default_unet_64.pt", "loss_cases.pt", "small_ae.pt", "small_clip_unet.pt"]:
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    print("default UNet sha256", sha)


if __name__ == "__main__":
    main()
