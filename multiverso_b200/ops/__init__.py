from .get_gemm import get_gemm, get_gemm_supported

__all__ = ["get_gemm", "get_gemm_supported"]
