"""Drop-in for an UNCHANGED reference checkout: put this directory (and the repo root) on PYTHONPATH and run the reference's own
launcher --

    PYTHONPATH=<repo>/avid_cma_b200/dropin_site:<repo> python main-avid.py cfg.yaml --multiprocessing-distributed --world-size 1 --rank 0

Python imports `sitecustomize` at start-up of EVERY interpreter, so the parent and the workers `mp.spawn` starts
(main-avid.py:78) all bind the reference's module names (models, criterions, datasets, utils.*) to avid_cma_b200 before the
script's first import (avid_cma_b200/dropin.py)."""
import os

if os.environ.get('AVID_B200_DROPIN', '1') != '0':
    from avid_cma_b200 import dropin
    dropin.install()
