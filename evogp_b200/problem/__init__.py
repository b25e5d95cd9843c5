from .base import BaseProblem  # noqa: F401
from .symbolic_regression import SymbolicRegression  # noqa: F401
from .classification import Classification  # noqa: F401
from .transformation import Transformation  # noqa: F401
