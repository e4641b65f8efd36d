"""Shim of pytorchvideo 0.1.5 (only what /root/reference/model/x3d.py imports)."""
__version__ = "0.1.5+shim"
