"""The north-star sentence "main-avid.py drops in unchanged", executed.

CPU (this container, where /root/reference is mounted): the UNMODIFIED reference launcher is run by path under the
`sitecustomize` redirect of avid_cma_b200/dropin_site.  Without a visible GPU the reference's own script stops at
main-avid.py:100 (`model.module.out_dim` on an unwrapped model: SURVEY.md Appendix C) -- everything before that line (argument
parsing, yaml, prep_environment, build_model, distribute_model_to_cuda) must have run on this package.

GPU (`-m gpu`; /root/reference does not exist on the GPU box): the same redirect under a vendored stub of the launcher's call
sequence (tests/dropin_stub/launcher_stub.py), started with --multiprocessing-distributed so that `mp.spawn` workers, NCCL,
DistributedDataParallel, the end-of-epoch meter synchronisation and the checkpoint writer all run through the redirect."""
import copy
import os
import socket
import subprocess
import sys

import pytest
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SITE = os.path.join(ROOT, "avid_cma_b200", "dropin_site")
REFERENCE = "/root/reference/main-avid.py"

CFG = {
    'resume': False, 'num_workers': 0, 'log2tb': False, 'seed': 0, 'test_freq': 1, 'print_freq': 2,
    'dataset': {'name': 'synthetic', 'num_samples': 48, 'batch_size': 8, 'video_clip_duration': 0.25, 'video_fps': 16., 'crop_size': 32,
                'audio_clip_duration': 0.4, 'audio_fps': 24000., 'spectrogram_fps': 100., 'n_fft': 64,
                'train': {'split': 'train', 'use_augmentation': True, 'drop_last': True, 'clips_per_video': 1}},
    'optimizer': {'name': 'adam', 'weight_decay': 1e-5, 'num_epochs': 2,
                  'lr': {'name': 'multistep', 'base_lr': 2e-4, 'gamma': 0.5, 'milestones': [1]}},
    'model': {'name': 'dropin', 'model_dir': None, 'arch': 'av_wrapper',
              'args': {'proj_dim': [512, 512, 128], 'video_backbone': 'R2Plus1D', 'video_backbone_args': {'depth': 18},
                       'audio_backbone': 'Conv2D', 'audio_backbone_args': {'depth': 10}}},
    'loss': {'name': 'AVID', 'args': {'num_data': 48, 'num_negatives': 16, 'momentum': 0.5, 'xModal_coeff': 1., 'wModal_coeff': 0.}},
}


def _cfg(tmp_path):
    cfg = copy.deepcopy(CFG)
    cfg['model']['model_dir'] = str(tmp_path)
    path = tmp_path / 'cfg.yaml'
    path.write_text(yaml.safe_dump(cfg))
    return str(path)


def _env():
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([SITE, ROOT] + ([env['PYTHONPATH']] if env.get('PYTHONPATH') else []))
    env['PYTHONDONTWRITEBYTECODE'] = '1'          # /root/reference is read-only
    return env


def test_redirect_binds_the_reference_module_names():
    code = ("import sys, models, criterions, datasets, utils.logger; from utils import main_utils, metrics_utils;"
            "print(models.__name__, criterions.__name__, datasets.__name__, main_utils.__name__, utils.logger.__name__, metrics_utils.__name__);"
            "assert models.av_wrapper and criterions.AVID and criterions.AVID_CMA and main_utils.build_model")
    out = subprocess.run([sys.executable, "-c", code], env=_env(), capture_output=True, text=True, timeout=300, cwd="/")
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ['avid_cma_b200.models', 'avid_cma_b200.criterions', 'avid_cma_b200.datasets', 'avid_cma_b200.utils.main_utils',
                                  'avid_cma_b200.utils.logger', 'avid_cma_b200.utils.metrics_utils']


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="the reference checkout is only mounted in the build container")
@pytest.mark.skipif(torch.cuda.is_available(), reason="GPU-less path of the reference launcher")
def test_unmodified_reference_launcher_runs_on_this_package_until_its_gpu_less_defect(tmp_path):
    cfg = _cfg(tmp_path)
    out = subprocess.run([sys.executable, REFERENCE, cfg, '--quiet'], env=_env(), capture_output=True, text=True, timeout=600,
                         cwd=os.path.dirname(REFERENCE))
    log = (tmp_path / 'dropin' / 'train.log').read_text()
    # prep_environment, build_model and distribute_model_to_cuda of THIS package ran under the reference's own main_worker
    assert 'backend: avid_cma_b200' in log and 'libavid_b200.so' in log
    assert '   Config   ' in log and '   Args   ' in log and '   Model   ' in log and '   Parameters   ' in log
    assert 'video_model.conv2x.0.spt_conv1.weight' in log and 'audio_proj.projection.4.bias' in log
    # ... up to main-avid.py:100, which needs a DataParallel / DDP wrap that a GPU-less host does not get (main_utils.py:99-100)
    assert out.returncode != 0 and "has no attribute 'module'" in out.stderr and 'main-avid.py", line 100' in out.stderr, out.stderr[-2000:]


@pytest.mark.gpu
def test_launcher_call_sequence_under_the_redirect_multiprocessing_distributed(tmp_path):
    cfg = _cfg(tmp_path)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    stub = os.path.join(ROOT, "tests", "dropin_stub", "launcher_stub.py")
    out = subprocess.run([sys.executable, stub, cfg, '--quiet', '--multiprocessing-distributed', '--world-size', '1', '--rank', '0',
                          '--dist-url', 'tcp://127.0.0.1:%d' % port, '--seed', '0'],
                         env=_env(), capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert out.returncode == 0, out.stderr[-4000:]
    run_dir = tmp_path / 'dropin'
    log = (run_dir / 'train.log').read_text()
    assert 'backend: avid_cma_b200' in log and ' Epoch 0 ' in log and ' Epoch 1 ' in log and 'train [1]' in log
    ck = torch.load(run_dir / 'checkpoint.pth.tar', weights_only=False)
    assert ck['epoch'] == 2 and len(ck['model']) == 267 and all(k.startswith('module.') for k in ck['model'])
    assert {'nce_average.view1_mem', 'nce_average.view2_mem', 'criterion.avg_exp_score'} <= set(ck['train_criterion'])
    assert tuple(ck['train_criterion']['nce_average.view1_mem'].shape) == (48, 128)
    assert float(ck['train_criterion']['criterion.avg_exp_score']) > 0
