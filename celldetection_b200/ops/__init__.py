from . import cpn
from . import conv
from . import boxes
from .boxes import filter_by_box_voting, box_votes, contours2boxes  # noqa: F401
from .cpn import *  # noqa: F401,F403
