from .neptune import *  # noqa: F401,F403
from .neptune import (NeptuneBase, NeptuneMinDelay, NeptuneMinDelayAndUtilization, NeptuneMinUtilization,  # noqa: F401
                      NeptuneWithEFTTCMinDelay, NeptuneWithEFTTCMinDelayAndUtilization,
                      NeptuneWithEFTTCMinUtilization)
from .neptune_step1 import (NeptuneStep1CPUBase, NeptuneStep1CPUMinDelay,  # noqa: F401
                            NeptuneStep1CPUMinDelayAndUtilization, NeptuneStep1CPUMinUtilization, NeptuneStepBase)
from .neptune_step2 import (NeptuneStep2Base, NeptuneStep2MinDelay, NeptuneStep2MinDelayAndUtilization,  # noqa: F401
                            NeptuneStep2MinUtilization)
