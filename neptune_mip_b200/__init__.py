"""neptune_mip_b200 -- B200-native solve path for NEPTUNE's function/request placement MIP.

Drop-in for the reference's `core` package on the solve path: `check_input`, `data_to_solver_input`,
`Data` and the solver classes (`NeptuneMinDelay`, `EfttcMinDelay`, ...) keep the reference's names,
signatures and output JSON; the work is done by hand-written sm_100a kernels behind the C ABI in
`include/neptune_b200.h` (`libneptune_b200.so`).  No CPU fallback.
"""
from .core.utils import Data, check_input, data_to_solver_input  # noqa: F401

__all__ = ["Data", "check_input", "data_to_solver_input"]
