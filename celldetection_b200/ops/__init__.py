from . import cpn
from . import conv
from .cpn import *  # noqa: F401,F403
