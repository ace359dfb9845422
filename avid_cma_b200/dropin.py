"""Run the reference's UNCHANGED `main-avid.py` on this package.

The reference resolves everything it builds BY NAME from top-level packages of its own tree (`import models`,
`import criterions`, `import datasets`, `from utils import main_utils`, `import utils.logger`: main-avid.py:20-21,
utils/main_utils.py:15,76,143,232).  `install()` binds those names to this package's mirrors in `sys.modules`, so the script's
own call sequence (main-avid.py:84-201) builds the B200 model / criterion / optimizer / loaders / checkpoints without a single
edit:

    python -m avid_cma_b200.dropin /path/to/AVID-CMA/main-avid.py cfg.yaml [--multiprocessing-distributed --world-size 1 --rank 0 ...]

With `mp.spawn` workers (`--multiprocessing-distributed`) the redirect has to exist in every spawned interpreter as well, and
the workers re-import the launcher by path: use the `sitecustomize.py` of avid_cma_b200/dropin_site instead, which installs the
redirect at interpreter start-up --

    PYTHONPATH=<repo>/avid_cma_b200/dropin_site:<repo> python /path/to/AVID-CMA/main-avid.py cfg.yaml --multiprocessing-distributed ...
"""
import os
import runpy
import sys

# reference module name -> mirror in this package
REDIRECTS = {
    'models': 'avid_cma_b200.models',
    'criterions': 'avid_cma_b200.criterions',
    'datasets': 'avid_cma_b200.datasets',
    'utils': 'avid_cma_b200.utils',
    'utils.main_utils': 'avid_cma_b200.utils.main_utils',
    'utils.logger': 'avid_cma_b200.utils.logger',
    'utils.metrics_utils': 'avid_cma_b200.utils.metrics_utils',
    'utils.distributed_utils': 'avid_cma_b200.utils.distributed_utils',
}


def install():
    """Bind the reference's top-level module names to this package (idempotent).  Returns the mapping that was installed."""
    import importlib
    done = {}
    for name, target in REDIRECTS.items():
        mod = importlib.import_module(target)
        sys.modules[name] = mod
        done[name] = mod
    # `from utils import main_utils` looks the attribute up on the package object
    pkg = sys.modules['utils']
    for name in ('main_utils', 'logger', 'metrics_utils', 'distributed_utils'):
        setattr(pkg, name, sys.modules['utils.' + name])
    os.environ['AVID_B200_DROPIN'] = '1'
    return done


def run(script, argv):
    """runpy the (unmodified) reference launcher `script` with `argv` under the redirect."""
    install()
    old = sys.argv
    sys.argv = [script] + list(argv)
    try:
        return runpy.run_path(script, run_name='__main__')
    finally:
        sys.argv = old


if __name__ == '__main__':
    if len(sys.argv) < 3:
        sys.exit('usage: python -m avid_cma_b200.dropin /path/to/main-avid.py cfg.yaml [reference flags]')
    run(sys.argv[1], sys.argv[2:])
