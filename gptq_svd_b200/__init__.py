"""B200-native (sm_100a) TruncGPTQ solve-and-quantize hot path.

Drop-in for the reference's `gptq_utils` names; see gptq_svd_b200/gptq_utils.py."""
from .gptq_utils import (HessianAccumulator, Quantizer, QuantizedLinear, SpectralFactors,  # noqa: F401
                         export_gptq, gptq_fwrd, gptq_quantize, log_quantization_error, pack_codes,
                         process_hessian_alt, spectral_solve)

from .frontends import Sketcher, process_hessian, process_sketch  # noqa: F401,E402

__all__ = ["Sketcher", "process_hessian", "process_sketch", "HessianAccumulator", "Quantizer", "QuantizedLinear", "SpectralFactors", "gptq_fwrd",
           "export_gptq", "gptq_quantize", "log_quantization_error", "pack_codes", "process_hessian_alt", "spectral_solve"]
