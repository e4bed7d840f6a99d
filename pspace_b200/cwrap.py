"""Same role as the reference's pspace/cwrap.py: `import pspace_b200.cwrap as uq` gives the wrapper classes."""
from .PSPACE import bind as _bind  # noqa: F401


def __getattr__(name):
    from . import PSPACE
    return getattr(PSPACE, name)
