"""B200-native ``tome`` package: same import surface as the reference's ``tome``
(tome/__init__.py:8-12) for the token-merging path -- ``tome.merge``, ``tome.patch``,
``tome.utils`` -- backed by hand-written sm_100a CUDA kernels (libtome_b200.so)."""
from . import merge, utils  # noqa: F401
from . import patch  # noqa: F401

__all__ = ["utils", "merge", "patch"]
