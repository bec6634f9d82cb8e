from . import cpn
from .cpn import *  # noqa: F401,F403
