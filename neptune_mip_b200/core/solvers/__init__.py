"""`from core.solvers import *` of the reference (`core/solvers/__init__.py:1-5`): the Neptune* and
Efttc* families.  VSVBP / Criticality / MCF are commented out in the reference and absent here."""
from .efttc import *  # noqa: F401,F403
from .efttc import (EfttcBase, EfttcMinDelay, EfttcMinDelayAndUtilization, EfttcMinUtilization,  # noqa: F401
                    EfttcStep1CPUBase, EfttcStep1CPUMinDelay, EfttcStep1CPUMinDelayAndUtilization,
                    EfttcStep1CPUMinUtilization, EfttcStepBase)
from .neptune import *  # noqa: F401,F403
from .neptune import (NeptuneBase, NeptuneMinDelay, NeptuneMinDelayAndUtilization, NeptuneMinUtilization,  # noqa: F401
                      NeptuneStep1CPUBase, NeptuneStep1CPUMinDelay, NeptuneStep1CPUMinDelayAndUtilization,
                      NeptuneStep1CPUMinUtilization, NeptuneStep2Base, NeptuneStep2MinDelay,
                      NeptuneStep2MinDelayAndUtilization, NeptuneStep2MinUtilization, NeptuneStepBase,
                      NeptuneWithEFTTCMinDelay, NeptuneWithEFTTCMinDelayAndUtilization,
                      NeptuneWithEFTTCMinUtilization)
from .output import convert_c_matrix, convert_x_matrix  # noqa: F401
from .solver import Solver  # noqa: F401
