"""lgar_c_b200 -- B200-native batched LGAR/CASAM engine (host-side Python mirror of the C ABI).

The product is ``liblgarb.so`` (hand-written sm_100a CUDA kernels + a C ABI, include/lgarb.h).  This
package only binds it with ctypes and mirrors the reference's BMI class (``BmiLGAR``, reference
include/bmi_lgar.hxx:30-184) for a column batch.  There is no CPU path: if the library or a CUDA
device is missing, construction/initialisation raises.
"""
from .bmi import BmiLGARBatch, LgarbError, build_library, library_path, OUTPUT_SCALARS, MAX_FRONTS, STATUS_FAIL_MASK

__all__ = ["BmiLGARBatch", "LgarbError", "build_library", "library_path", "OUTPUT_SCALARS", "MAX_FRONTS", "STATUS_FAIL_MASK"]
