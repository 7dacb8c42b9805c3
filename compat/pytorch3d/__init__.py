"""Import shim: the few PyTorch3D names the reference's scripts use once its models are replaced by forge_b200."""
__version__ = "0.7.0+forge_b200.compat"
