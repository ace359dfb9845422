"""avid_cma_b200 -- B200-native implementation of the AVID / AVID-CMA training hot path.

Drop-in surface (same names / kwargs / state_dict keys as facebookresearch/AVID-CMA):
    avid_cma_b200.models.av_wrapper(...)            <- models/av_wrapper.py:64
    avid_cma_b200.criterions.AVID(...), AVID_CMA(...)   <- criterions/avid.py:145, criterions/avid_cma.py:245
Every device computation goes through libavid_b200.so (include/avid_b200.h); there is no CPU path.
"""
__version__ = "0.1.0"
