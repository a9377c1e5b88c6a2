from .sampler import Sampler  # noqa: F401
from .container import SamplerContainer  # noqa: F401
from .slice import Slice  # noqa: F401
from .transform import Transform  # noqa: F401
from .function import Function, FunctionDerivative  # noqa: F401
