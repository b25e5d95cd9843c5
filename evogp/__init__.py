"""Import-path shim: existing EvoGP scripts (`from evogp.tree import Forest`, `evogp.algorithm`,
`evogp.problem`, `evogp.pipeline`) run unchanged on the B200-native implementation."""
import importlib
import sys

import evogp_b200 as _impl

__version__ = _impl.__version__
for _sub in ("tree", "algorithm", "problem", "pipeline"):
    _mod = importlib.import_module(f"evogp_b200.{_sub}")
    sys.modules[f"{__name__}.{_sub}"] = _mod
    setattr(sys.modules[__name__], _sub, _mod)
