"""Mirror of the reference's `core` package (`core/__init__.py`): the input adapter and `Data`."""
from .utils import Data, check_input, data_to_solver_input  # noqa: F401
