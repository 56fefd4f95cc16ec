from . import numerize as _m  # noqa: F401
from .numerize import numerize  # noqa: F401
