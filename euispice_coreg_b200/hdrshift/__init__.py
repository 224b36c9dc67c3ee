from .alignment import Alignment  # noqa: F401
from .AlignmentResults import AlignmentResults  # noqa: F401
