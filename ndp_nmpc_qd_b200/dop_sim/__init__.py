from .mul_quadrotors import MulQuadrotors  # noqa: F401
