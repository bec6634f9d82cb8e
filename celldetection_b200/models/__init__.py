from . import cpn as _cpn
from .cpn import *  # noqa: F401,F403  (CPN and every Cpn<Encoder><Decoder> class of models.graph.ARCHS)

__all__ = list(_cpn.__all__)
