"""Drop-in counterparts of the reference's `model/` package (x3d, change_decoder, trainer, utils)."""
