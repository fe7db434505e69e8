"""GPU, harness level: the reference trainer's loop shapes driven against the drop-in
with a synthetic loader of dict batches {"image", "mask", "original_dims"} (what `PetSegmentationDataset` yields,
Our_UNet/src/train.py:313-320).

The reference tree is not on the GPU box, so its functions are restated here line for line in what they DO to the model
and loss (the parts that matter for a drop-in): `train_one_epoch` (train.py:592-680: H2D copy, zero_grad,
[autocast +] forward, loss, [scaled] backward, step, loss.item()), `validate` (train.py:510-589: eval, no_grad, argmax,
per-class Dice with nine .item() syncs), `save_checkpoint` (train.py:683-739: the dict layout), resume (train.py:888-902)
and `evaluate.load_model` (evaluate.py:103-147: constructor kwargs, `checkpoint["model_state_dict"]`, eval mode).
tests/test_dropin.py runs the UNMODIFIED `src/train.py` against the drop-in on the CPU box as far as imports and
construction go; this file is the part that needs a GPU.
"""
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

KW = dict(in_channels=3, num_classes=3, n_stages=6, features_per_stage=[32, 64, 128, 256, 512, 512],
          kernel_sizes=[[3, 3]] * 6, strides=[[1, 1]] + [[2, 2]] * 5, n_conv_per_stage=[2] * 6,
          n_conv_per_stage_decoder=[2] * 5, conv_bias=True, norm_op=nn.InstanceNorm2d,
          norm_op_kwargs={"eps": 1e-5, "affine": True}, dropout_op=None, nonlin=nn.LeakyReLU,
          nonlin_kwargs={"inplace": True}, encoder_dropout_rates=[0.0, 0.0, 0.1, 0.2, 0.3, 0.3],
          decoder_dropout_rates=[0.3, 0.2, 0.2, 0.1, 0.0])  # train.py:776-795 == evaluate.py:115-132


def loader(n_batches, batch, size, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        img = torch.randn(batch, 3, size, size, generator=g)
        # learnable structure: the class is a function of the image
        mask = (img[:, 0] > 0.3).long() + (img[:, 1] > 0.8).long()
        mask[torch.rand(batch, size, size, generator=g) < 0.05] = 255
        out.append({"image": img, "mask": mask, "original_dims": [(size, size)] * batch})
    return out


def train_one_epoch(model, train_loader, optimizer, loss_function, device, scaler=None):  # train.py:592-680
    model.train()
    epoch_loss = 0.0
    for batch in train_loader:
        images = batch["image"].to(device)
        masks = batch["mask"].to(device)
        optimizer.zero_grad()
        if scaler is not None:
            with torch.autocast("cuda", dtype=torch.float16):
                outputs = model(images)
                loss = loss_function(outputs, masks)
            scaler.scale(loss).backward()
            scaler.step(optimizer)
            scaler.update()
        else:
            outputs = model(images)
            loss = loss_function(outputs, masks)
            loss.backward()
            optimizer.step()
        epoch_loss += loss.item()
    return epoch_loss / len(train_loader)


def validate(model, val_loader, loss_function, device, ignore_label=255):  # train.py:510-589
    model.eval()
    val_loss = 0.0
    dice_scores = {"background": 0.0, "cat": 0.0, "dog": 0.0, "mean_foreground": 0.0}
    with torch.no_grad():
        for batch in val_loader:
            images = batch["image"].to(device)
            masks = batch["mask"].to(device)
            outputs = model(images)
            loss = loss_function(outputs, masks)
            val_loss += loss.item()
            preds = torch.argmax(outputs, dim=1)
            for cls_idx, cls_name in [(0, "background"), (1, "cat"), (2, "dog")]:
                pred_cls = (preds == cls_idx).float()
                mask_cls = (masks == cls_idx).float()
                ignore_mask = (masks != ignore_label).float()
                pred_cls = pred_cls * ignore_mask
                mask_cls = mask_cls * ignore_mask
                intersection = (pred_cls * mask_cls).sum()
                union = pred_cls.sum() + mask_cls.sum()
                dice = (2.0 * intersection) / (union + 1e-5) if union > 0 else torch.tensor(1.0, device=device)
                dice_scores[cls_name] += dice.item()
    for k in dice_scores:
        dice_scores[k] /= len(val_loader)
    dice_scores["mean_foreground"] = (dice_scores["cat"] + dice_scores["dog"]) / 2.0
    return val_loss / len(val_loader), dice_scores


def save_checkpoint(model, optimizer, scheduler, epoch, best_dice, output_dir):  # train.py:683-739
    ckpt = {"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
            "scheduler_state_dict": scheduler.state_dict(), "best_dice": best_dice, "config": {"n_stages": 8}}
    path = os.path.join(output_dir, f"checkpoint_epoch_{epoch}.pth")
    torch.save(ckpt, path)
    return path


def make_optimizer(model, fused):
    kw = dict(lr=0.005, weight_decay=1e-4, momentum=0.99, nesterov=True)  # train.py:445-451
    if fused:
        from unet_implementations_b200.optim import FusedSGD
        return FusedSGD(model.parameters(), model=model, **kw)
    return torch.optim.SGD(model.parameters(), **kw)


@pytest.mark.parametrize("fused", [False, True])
def test_train_validate_checkpoint_resume_evaluate(tmp_path, fused):
    from unet_implementations_b200 import metrics
    from unet_implementations_b200.models.losses import SimpleLoss
    from unet_implementations_b200.models.unet import UNet
    device = torch.device("cuda")
    torch.manual_seed(1234)
    model = UNet(**KW).to(device)
    optimizer = make_optimizer(model, fused)
    epochs = 3
    scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lambda e: (1 - e / epochs) ** 0.9)  # train.py:466-475
    loss_function = SimpleLoss(weight_dice=1.0, weight_ce=1.0, ignore_index=255, smooth=1e-5, class_weights=None,
                               dynamic_weights=True)  # train.py:862-869
    train_loader, val_loader = loader(6, 4, 128, 1), loader(2, 4, 128, 2)
    v0, _ = validate(model, val_loader, loss_function, device)
    losses = []
    for epoch in range(2):
        losses.append(train_one_epoch(model, train_loader, optimizer, loss_function, device))
        scheduler.step()
    v1, dice = validate(model, val_loader, loss_function, device)
    assert all(l == l for l in losses) and losses[1] < losses[0], losses    # it trains
    assert v1 < v0, (v0, v1)
    assert all(0.0 <= dice[k] <= 1.0 for k in dice)
    # the fused validation metric gives the same Dice as validate()'s nine-sync arithmetic
    with torch.no_grad():
        b = val_loader[0]
        out = model(b["image"].to(device))
        _, counts = metrics.argmax_counts(out, b["mask"].to(device))
        d = metrics.dice_from_counts(counts)
        preds = torch.argmax(out, 1)
        m = b["mask"].to(device)
        for c in range(3):
            inter = ((preds == c) & (m == c)).sum().float()
            union = ((preds == c) & (m != 255)).sum().float() + (m == c).sum().float()
            ref = (2 * inter / (union + 1e-5)) if union > 0 else torch.tensor(1.0)
            assert abs(float(d[c]) - float(ref)) < 1e-6
    # checkpoint -> resume (train.py:888-902) -> the continued run equals the uninterrupted one, bit for bit
    path = save_checkpoint(model, optimizer, scheduler, 2, dice["mean_foreground"], str(tmp_path))
    torch.manual_seed(7)
    cont = train_one_epoch(model, train_loader, optimizer, loss_function, device)
    ref_state = {k: v.clone() for k, v in model.state_dict().items()}

    torch.manual_seed(999)  # different init: everything must come from the checkpoint
    model2 = UNet(**KW).to(device)
    optimizer2 = make_optimizer(model2, fused)
    scheduler2 = torch.optim.lr_scheduler.LambdaLR(optimizer2, lambda e: (1 - e / epochs) ** 0.9)
    ckpt = torch.load(path, map_location=device, weights_only=False)
    assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "best_dice", "config"}
    assert len(ckpt["model_state_dict"]) == 90 and all(v.dtype == torch.float32 for v in ckpt["model_state_dict"].values())
    model2.load_state_dict(ckpt["model_state_dict"])
    optimizer2.load_state_dict(ckpt["optimizer_state_dict"])
    scheduler2.load_state_dict(ckpt["scheduler_state_dict"])
    torch.manual_seed(7)
    cont2 = train_one_epoch(model2, train_loader, optimizer2, loss_function, device)
    assert cont2 == cont
    for k, v in model2.state_dict().items():
        assert torch.equal(v, ref_state[k]), k
    # evaluate.load_model (evaluate.py:103-147): bare state_dict or wrapped, eval mode, argmax on the logits
    model3 = UNet(**KW)
    model3.load_state_dict(ckpt["model_state_dict"])
    model3 = model3.to(device).eval()
    model.load_state_dict(ckpt["model_state_dict"])
    model.eval()
    with torch.no_grad():
        x = val_loader[1]["image"].to(device)
        assert torch.equal(torch.argmax(model3(x), 1), torch.argmax(model(x), 1))


def test_amp_fp16_gradscaler_epoch_with_the_fused_optimizer():
    """train.py:638-651: autocast + GradScaler around the drop-in; the scaler unscales the flat gradient buffer in place."""
    from unet_implementations_b200.models.losses import SimpleLoss
    from unet_implementations_b200.models.unet import UNet
    device = torch.device("cuda")
    torch.manual_seed(1234)
    model = UNet(**KW).to(device)
    optimizer = make_optimizer(model, True)
    scaler = torch.amp.GradScaler("cuda")
    l0 = train_one_epoch(model, loader(4, 2, 64, 3), optimizer, SimpleLoss(), device, scaler)
    l1 = train_one_epoch(model, loader(4, 2, 64, 3), optimizer, SimpleLoss(), device, scaler)
    assert l0 == l0 and l1 < l0
