"""Mirror of the reference's `models` package (Our_UNet/models/): `unet.UNet`, `losses.SimpleLoss`."""
