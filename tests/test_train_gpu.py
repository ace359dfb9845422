"""End-to-end run of the launcher (main_avid.py: the reference's main-avid.py sequence on the B200 path) on one GPU: two epochs
on a tiny synthetic dataset, reference-layout checkpoint, resume, and an AVID-CMA epoch with positive re-mining."""
import copy
import os

import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

CFG = {
    'resume': False, 'num_workers': 0, 'log2tb': False, 'seed': 0, 'test_freq': 1, 'print_freq': 2,
    'dataset': {'name': 'synthetic', 'num_samples': 24, 'batch_size': 4, 'video_clip_duration': 0.25, 'video_fps': 16., 'crop_size': 64,
                'audio_clip_duration': 0.6, 'audio_fps': 24000., 'spectrogram_fps': 100., 'n_fft': 128,
                'train': {'split': 'train', 'use_augmentation': True, 'drop_last': True, 'clips_per_video': 1}},
    'optimizer': {'name': 'adam', 'weight_decay': 1e-5, 'num_epochs': 2, 'lr': {'name': 'multistep', 'base_lr': 2e-4, 'gamma': 0.5, 'milestones': [1]}},
    'model': {'name': 'tiny', 'model_dir': None, 'arch': 'av_wrapper',
              'args': {'proj_dim': [512, 512, 128], 'video_backbone': 'R2Plus1D', 'video_backbone_args': {'depth': 18},
                       'audio_backbone': 'Conv2D', 'audio_backbone_args': {'depth': 10}}},
    'loss': {'name': 'AVID', 'args': {'num_data': 24, 'num_negatives': 64, 'momentum': 0.5, 'xModal_coeff': 1., 'wModal_coeff': 0.}},
}


def _write(cfg, path):
    with open(path, 'w') as f:
        yaml.safe_dump(cfg, f)
    return str(path)


def test_launcher_trains_checkpoints_and_resumes(tmp_path, monkeypatch):
    import main_avid
    monkeypatch.setenv("AVID_MATH", "bf16x3")
    cfg = copy.deepcopy(CFG)
    cfg['model']['model_dir'] = str(tmp_path)
    main_avid.main([_write(cfg, tmp_path / 'cfg.yaml'), '--quiet', '--seed', '0'])
    run_dir = tmp_path / 'tiny'
    log = open(run_dir / 'train.log').read()
    assert 'Epoch 0' in log and 'Epoch 1' in log and 'train [1][6/6]' in log and 'LR: [0.0001]' in log
    files = sorted(os.listdir(run_dir))
    assert files == ['checkpoint-ep1.pth.tar', 'checkpoint.pth.tar', 'train.log']
    ck = torch.load(run_dir / 'checkpoint.pth.tar', weights_only=False)
    assert ck['epoch'] == 2 and set(ck) == {'epoch', 'model', 'optimizer', 'train_criterion'}
    # the reference's 267 model keys behind the DataParallel 'module.' prefix, its criterion keys, Adam's state layout
    assert len(ck['model']) == 267 and all(k.startswith('module.') for k in ck['model'])
    assert ck['model']['module.video_model.conv2x.0.spt_conv1.weight'].shape == (64, 64, 1, 3, 3)
    assert set(ck['train_criterion']) == {'nce_average.view1_mem', 'nce_average.view2_mem', 'criterion.avg_exp_score', 'nce_average.sampler_state'}   # the reference's keys + our Philox stream position (the reference restores with strict=False)
    assert ck['train_criterion']['nce_average.view1_mem'].shape == (24, 128)
    st = ck['optimizer']['state']
    assert len(st) == len(ck['optimizer']['param_groups'][0]['params']) and {'step', 'exp_avg', 'exp_avg_sq'} <= set(st[0])
    assert all(torch.isfinite(v).all() for v in ck['model'].values() if v.is_floating_point())
    # resume: one more epoch from the checkpoint, learning rate carried over (0.5 * base after milestone 1)
    cfg['resume'] = True
    cfg['optimizer']['num_epochs'] = 3
    main_avid.main([_write(cfg, tmp_path / 'cfg2.yaml'), '--quiet'])
    log2 = open(run_dir / 'train.log').read()
    assert "Checkpoint loaded" in log2 and '(epoch 2)' in log2 and 'Epoch 2' in log2 and 'Epoch 0' not in log2 and 'LR: [0.0001]' in log2
    ck2 = torch.load(run_dir / 'checkpoint.pth.tar', weights_only=False)
    assert ck2['epoch'] == 3 and int(ck2['optimizer']['state'][0]['step']) == 18
    moved = (ck2['model']['module.audio_model.block4.conv2.weight'] - ck['model']['module.audio_model.block4.conv2.weight']).abs().max()
    assert float(moved) > 0


def test_launcher_avid_cma_epoch(tmp_path, monkeypatch):
    import main_avid
    monkeypatch.setenv("AVID_MATH", "bf16x3")
    cfg = copy.deepcopy(CFG)
    cfg['model']['model_dir'] = str(tmp_path)
    cfg['model']['name'] = 'tiny-cma'
    cfg['dataset']['num_samples'] = 200
    cfg['dataset']['batch_size'] = 8
    cfg['optimizer']['num_epochs'] = 2
    cfg['print_freq'] = 100
    cfg['loss'] = {'name': 'AVID_CMA', 'args': {'num_data': 200, 'num_negatives': 64, 'num_negatives_within': 16, 'momentum': 0.5,
                                                 'xModalInstCoeff': 1., 'wModalInstCoeff': 0., 'xModalPosCoeff': 0., 'wModalPosCoeff': 1.,
                                                 'sampling_args': {'type': 'consensus', 'pos_k': 8}, 'resample_freq': 1}}
    main_avid.main([_write(cfg, tmp_path / 'cfg.yaml'), '--quiet'])
    ck = torch.load(tmp_path / 'tiny-cma' / 'checkpoint.pth.tar', weights_only=False)
    ps = ck['train_criterion']['nce_average.positive_set']
    assert ps.shape == (200, 8) and int(ps.min()) >= 0 and int(ps.max()) < 200
    assert bool((ps[:, 1:] > ps[:, :-1]).all())                     # sorted ascending (avid_cma.py:70)
    assert not bool((ps == torch.arange(200).view(-1, 1)).any())     # never the instance itself
