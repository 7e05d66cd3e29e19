from .get_gemm import get_gemm, get_gemm_supported
from .ps_linear import PSLinear, ps_linear

__all__ = ["get_gemm", "get_gemm_supported", "ps_linear", "PSLinear"]
