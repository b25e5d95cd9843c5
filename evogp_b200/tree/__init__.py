from .. import _native

_native.load_ops()  # the reference imports its extension here for the same side effect (tree/__init__.py:2)

from .utils import MAX_STACK, MAX_FULL_DEPTH, NType, Func, FUNCS_NAMES, randint  # noqa: E402,F401
from .descriptor import GenerateDescriptor  # noqa: E402,F401
from .tree import Tree  # noqa: E402,F401
from .forest import Forest  # noqa: E402,F401
