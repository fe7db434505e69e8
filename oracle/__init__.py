"""Test infrastructure only (see unet_oracle.py's header): nothing in the product imports this package."""
