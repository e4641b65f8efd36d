"""Shim of fvcore 0.1.5.post20221221 (only SqueezeExcitation is needed)."""
