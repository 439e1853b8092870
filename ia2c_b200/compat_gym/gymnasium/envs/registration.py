from ..core import register  # noqa: F401
