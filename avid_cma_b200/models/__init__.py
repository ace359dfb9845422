from .video import *        # noqa: F401,F403
from .video import R2Plus1D  # noqa: F401
from .audio import *        # noqa: F401,F403
from .av_wrapper import *   # noqa: F401,F403
from .av_wrapper import AV_Wrapper, Head   # noqa: F401
