from .collectives import aggregate, symm_tensor

__all__ = ["aggregate", "symm_tensor"]
