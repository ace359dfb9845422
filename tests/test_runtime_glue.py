"""Host-side glue of the training loop: meters, logger, optimizer / scheduler builder, synthetic loader,
checkpoint manager.  CPU only.  Where the reference tree is mounted (/root/reference, this container only) the meters are
additionally compared with the reference's own classes on the same update sequence."""
import argparse
import importlib.util
import os
import sys

import pytest
import torch
import torch.nn as nn

from avid_cma_b200.utils import logger as L
from avid_cma_b200.utils import main_utils as MU
from avid_cma_b200.utils import metrics_utils as M

REF = "/root/reference"


def _ref_module(rel, name):
    if not os.path.isfile(os.path.join(REF, rel)):
        return None
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(mod)
    return mod


def test_average_meter_plain_and_windowed():
    m = M.AverageMeter('Loss', ':.3e')
    for v, n in [(2.0, 4), (4.0, 4), (1.0, 8)]:
        m.update(v, n)
    assert m.val == 1.0 and m.count == 16 and abs(m.avg - (8 + 16 + 8) / 16) < 1e-12
    assert str(m) == 'Loss 1.000e+00 (2.000e+00)'
    w = M.AverageMeter('Time', ':6.3f', window_size=2)
    for v in (1.0, 2.0, 6.0):
        w.update(v)
    assert w.avg == 4.0 and w.count == 2 and str(w) == 'Time  6.000 ( 4.000)'
    ref = _ref_module("utils/metrics_utils.py", "ref_metrics_utils")
    if ref is not None:
        for kw in (dict(fmt=':.3e'), dict(fmt=':6.3f', window_size=3)):
            a, b = M.AverageMeter('X', **kw), ref.AverageMeter('X', **kw)
            g = torch.Generator().manual_seed(0)
            for _ in range(10):
                v, n = float(torch.rand((), generator=g)), int(torch.randint(1, 5, (), generator=g))
                a.update(v, n); b.update(v, n)
                assert str(a) == str(b) and a.avg == b.avg and a.count == b.count
        out, tgt = torch.randn(16, 10, generator=torch.Generator().manual_seed(1)), torch.arange(16) % 10
        # (the reference's accuracy() fails for k > 1 on torch >= 1.7: .view on a non-contiguous slice, metrics_utils.py:24)
        assert torch.equal(M.accuracy(out, tgt, (1,))[0], ref.accuracy(out, tgt, (1,))[0])
    out, tgt = torch.randn(16, 10, generator=torch.Generator().manual_seed(1)), torch.arange(16) % 10
    top1, top5 = M.accuracy(out, tgt, (1, 5))
    assert float(top1) == 100.0 * float((out.argmax(1) == tgt).float().mean())
    assert float(top5) == 100.0 * float((out.topk(5, 1).indices == tgt[:, None]).any(1).float().mean())


def test_logger_and_progress_meter(tmp_path, capsys):
    fn = str(tmp_path / "train.log")
    lg = L.Logger(quiet=True, log_fn=fn, rank=0, prefix="p")
    lg.add_line("hello")
    L.Logger(quiet=True, log_fn=str(tmp_path / "other.log"), rank=1).add_line("silent")
    assert open(fn).read() == "p | hello\n" and not os.path.exists(str(tmp_path / "other.log"))
    meters = [M.AverageMeter('Time', ':6.3f'), M.AverageMeter('Loss', ':.3e')]
    meters[0].update(0.5); meters[1].update(3.0)
    pm = L.ProgressMeter(120, meters, phase='train', epoch=3, logger=L.Logger(quiet=False, rank=0))
    assert pm.batch_fmtstr.format(7) == '[3][  7/120]'
    assert pm.progress.meters is meters                 # the attribute main-avid.py:196 -> logger.py:74 dereferences
    pm.display(7)
    line = capsys.readouterr().out.strip()
    assert line.endswith('train [3][  7/120]\tTime  0.500 ( 0.500)\tLoss 3.000e+00 (3.000e+00)')


def _sync_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    m = M.AverageMeter('Loss', ':.3e')
    m.update(float(rank + 1))
    pm = L.ProgressMeter(10, [m], phase='train', epoch=0)
    pm.synchronize_meters(None)
    q.put((rank, m.avg))
    dist.destroy_process_group()


def test_synchronize_meters_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    ps = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    assert got == {0: 1.5, 1: 1.5}


def test_build_optimizer_and_schedule():
    net = nn.Linear(4, 4)
    cfg = {'name': 'adam', 'weight_decay': 1e-5, 'lr': {'base_lr': 2e-4, 'gamma': 0.1, 'milestones': [1, 3]}}
    opt, sched = MU.build_optimizer(net.parameters(), cfg)
    assert isinstance(opt, torch.optim.Adam) and opt.defaults['betas'] == (0.9, 0.999) or list(opt.defaults['betas']) == [0.9, 0.999]
    lrs = []
    for _ in range(4):
        lrs.append(opt.param_groups[0]['lr'])
        opt.step(); sched.step()
    assert [round(x / 2e-4, 6) for x in lrs] == [1.0, 0.1, 0.1, 0.01]
    sgd, _ = MU.build_optimizer(net.parameters(), {'name': 'sgd', 'momentum': 0.9, 'weight_decay': 0., 'nesterov': True,
                                                   'lr': {'base_lr': 0.1, 'gamma': 1., 'milestones': []}})
    assert isinstance(sgd, torch.optim.SGD)
    with pytest.raises(ValueError):
        MU.build_optimizer(net.parameters(), {'name': 'lamb', 'lr': {'base_lr': 1, 'gamma': 1, 'milestones': []}})


def test_synthetic_loader_has_reference_batch_contract():
    db_cfg = {'name': 'synthetic', 'num_samples': 10, 'batch_size': 4, 'video_clip_duration': 0.5, 'video_fps': 16., 'crop_size': 32,
              'audio_clip_duration': 2., 'audio_fps': 24000., 'spectrogram_fps': 100., 'n_fft': 512,
              'train': {'split': 'train', 'use_augmentation': True, 'drop_last': True, 'clips_per_video': 3}}
    loader = MU.build_dataloader(db_cfg, db_cfg['train'], num_workers=0, distributed=False)
    assert len(loader.dataset) == 30 and len(loader) == 7
    batch = next(iter(loader))
    assert batch['frames'].shape == (4, 3, 8, 32, 32) and batch['audio'].shape == (4, 1, 200, 257)
    assert batch['index'].dtype == torch.int64 and int(batch['index'].max()) < 10       # video_db.py:98: index % num_samples
    a, b = loader.dataset[13], loader.dataset[13]
    assert torch.equal(a['frames'], b['frames']) and a['index'] == 3
    with pytest.raises(ValueError):
        MU.build_dataloader(dict(db_cfg, name='kinetics'), db_cfg['train'], 0, False)


class _FakeBank(nn.Module):
    def __init__(self, sharded):
        super().__init__()
        self.sharded = sharded
        self.register_buffer('view1_mem', torch.ones(2, 4))
        self.register_buffer('view2_mem', torch.ones(2, 4) * 2)

    def full_banks(self):
        return torch.arange(16.).view(4, 4), -torch.arange(16.).view(4, 4)


class _FakeCriterion(nn.Module):
    def __init__(self, sharded=False):
        super().__init__()
        self.nce_average = _FakeBank(sharded)
        self.criterion = nn.Module()
        self.criterion.register_buffer('avg_exp_score', torch.tensor([2.5]))


def test_checkpoint_manager_round_trip(tmp_path):
    d = str(tmp_path)
    model = nn.DataParallel(nn.Linear(3, 2)) if False else nn.Sequential(nn.Linear(3, 2))
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    model(torch.randn(5, 3)).sum().backward(); opt.step()
    crit = _FakeCriterion(sharded=True)
    ck = MU.CheckpointManager(d, rank=0)
    assert not ck.checkpoint_exists(last=True)
    ck.save(7, model=model, optimizer=opt, train_criterion=crit)
    ck.save(3, model=model, train_criterion=crit, optimizer=opt, filename='checkpoint-ep3.pth.tar')
    MU.CheckpointManager(d, rank=1).save(9, model=model, filename='rank1.pth.tar')       # non-zero ranks never write
    assert sorted(os.listdir(d)) == ['checkpoint-ep3.pth.tar', 'checkpoint.pth.tar']
    raw = torch.load(ck.last_checkpoint_fn(), weights_only=False)
    assert set(raw) == {'epoch', 'model', 'optimizer', 'train_criterion'} and raw['epoch'] == 7
    # the sharded bank was gathered into the reference's full layout
    assert raw['train_criterion']['nce_average.view1_mem'].shape == (4, 4)
    assert set(raw['train_criterion']) == {'nce_average.view1_mem', 'nce_average.view2_mem', 'criterion.avg_exp_score'}
    model2 = nn.Sequential(nn.Linear(3, 2)); opt2 = torch.optim.Adam(model2.parameters(), lr=1e-3)
    crit2 = _FakeCriterion(sharded=False); crit2.nce_average.view1_mem = torch.zeros(4, 4); crit2.nce_average.view2_mem = torch.zeros(4, 4)
    assert ck.restore(restore_last=True, model=model2, optimizer=opt2, train_criterion=crit2) == 7
    assert torch.equal(model2[0].weight, model[0].weight) and opt2.state_dict()['state'][0]['step'] == opt.state_dict()['state'][0]['step']
    assert torch.equal(crit2.nce_average.view1_mem, torch.arange(16.).view(4, 4))
    ck.save(8, eval_metric=0.5, model=model)
    assert os.path.isfile(ck.best_checkpoint_fn()) and ck.checkpoint_exists(best=True)


def test_launcher_cli_matches_reference_flags():
    import main_avid
    a = main_avid.get_parser().parse_args(['cfg.yaml', '--multiprocessing-distributed', '--world-size', '1', '--rank', '0', '--seed', '3', '--quiet'])
    assert a.cfg == 'cfg.yaml' and a.multiprocessing_distributed and a.world_size == 1 and a.rank == 0 and a.seed == 3 and a.quiet
    assert a.dist_backend == 'nccl' and a.gpu is None
    args = argparse.Namespace(distributed=False, rank=-1, dist_url='tcp://127.0.0.1:1', multiprocessing_distributed=False)
    assert MU.initialize_distributed_backend(args, 0).rank == 0


def test_video_prep_host_logic_draws_like_the_oracle():
    """datasets.gpu_preprocessing.VideoPrep_MSC_CJ.draw consumes the `random` stream in the reference's order (the oracle's draw_params is
    pinned to the reference's goldens in tests/test_oracle_video.py)."""
    import random
    from avid_cma_b200.datasets.gpu_preprocessing import VideoPrep_MSC_CJ
    from oracle import video as V
    prep = VideoPrep_MSC_CJ(crop=(64, 64))
    for seed, (w, h) in enumerate([(128, 96), (100, 150), (340, 256), (20, 400), (400, 20)]):
        random.seed(seed)
        a = prep.draw(w, h)
        state = random.getstate()
        random.seed(seed)
        assert a == V.draw_params(w, h) and random.getstate() == state
    prep = VideoPrep_MSC_CJ(crop=(64, 64), color=(0.4, 0., 0.4, 0.), min_area=0.5)
    random.seed(7)
    a = prep.draw(128, 96)
    random.seed(7)
    assert a == V.draw_params(128, 96, min_area=0.5, color=(0.4, 0., 0.4, 0.)) and [o[0] for o in a['ops']] in (['brightness', 'saturation'], ['saturation', 'brightness'])
