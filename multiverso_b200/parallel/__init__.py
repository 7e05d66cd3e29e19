from .collectives import aggregate

__all__ = ["aggregate"]
