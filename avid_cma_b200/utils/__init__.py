"""Host-side helpers mirroring the reference's utils/ package for the training path (logger, metrics_utils, main_utils,
distributed_utils).  Submodules are imported on demand (`from avid_cma_b200.utils import main_utils`)."""
