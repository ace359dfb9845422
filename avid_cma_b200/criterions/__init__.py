from .avid import *       # noqa: F401,F403
from .avid_cma import *   # noqa: F401,F403
from .avid import AVIDSimilarityMemoryBank   # noqa: F401
from .avid_cma import AVIDSimilarityPositiveExpansion   # noqa: F401
from .nce import NCECriterion   # noqa: F401
