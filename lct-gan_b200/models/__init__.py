"""Drop-in `models` package: LCT-GAN generator and discriminators on lctgan sm_100a kernels."""
