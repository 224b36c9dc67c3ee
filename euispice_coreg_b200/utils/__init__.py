from . import Util  # noqa: F401
from .Util import AlignCommonUtil, AlignEUIUtil, AlignSpiceUtil  # noqa: F401
