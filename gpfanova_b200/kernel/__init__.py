from .kernel import Kernel, OFFSET  # noqa: F401
from .rbf import RBF  # noqa: F401
from .white import White  # noqa: F401
