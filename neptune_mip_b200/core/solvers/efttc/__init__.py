from .efttc import EfttcBase, EfttcMinDelay, EfttcMinDelayAndUtilization, EfttcMinUtilization  # noqa: F401
from .efttc_step1 import (EfttcStep1CPUBase, EfttcStep1CPUMinDelay,  # noqa: F401
                          EfttcStep1CPUMinDelayAndUtilization, EfttcStep1CPUMinUtilization, EfttcStepBase)
