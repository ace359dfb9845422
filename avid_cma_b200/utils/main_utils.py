"""Run-time glue with the call signatures of the reference's utils/main_utils.py, so that main-avid.py (or this package's
main_avid.py, which follows the same sequence) builds the B200 model, criterion, optimizer, loaders and checkpoints:

    initialize_distributed_backend  main_utils.py:18-31      build_criterion     main_utils.py:231-237
    prep_environment                main_utils.py:34-72      build_optimizer     main_utils.py:240-259
    build_model                     main_utils.py:75-95      CheckpointManager   main_utils.py:262-315
    distribute_model_to_cuda        main_utils.py:98-138     save_checkpoint / prep_output_folder / parameter_description
    build_dataloaders / build_dataloader  main_utils.py:141-228

Differences (all documented in INTEGRATION.md): models / criterions resolve to avid_cma_b200.{models,criterions}; 'adam'
builds the fused multi-tensor avid_cma_b200.optim.Adam (same update rule and state_dict layout as torch.optim.Adam);
dataset name 'synthetic' builds datasets.SyntheticAV (the PyAV / librosa datasets are out of scope); a row-sharded memory
bank is gathered into the reference's full (N,128) layout when a checkpoint is written; TensorBoard is optional.
"""
import datetime
import os
import shutil

import numpy as np
import torch
import torch.distributed as dist

from .logger import Logger


def initialize_distributed_backend(args, ngpus_per_node):
    if args.distributed:
        if args.dist_url == "env://" and args.rank == -1:
            args.rank = int(os.environ["RANK"])
        if args.multiprocessing_distributed:
            args.rank = args.rank * ngpus_per_node + args.gpu      # global rank of this process
        dist.init_process_group(backend=args.dist_backend, init_method=args.dist_url, world_size=args.world_size, rank=args.rank)
    if args.rank == -1:
        args.rank = 0
    return args


def _log_tree(logger, d, indent=''):
    for k, v in d.items():
        if isinstance(v, dict):
            logger.add_line("{}{}".format(indent, k))
            _log_tree(logger, v, '  ' + indent)
        else:
            logger.add_line("{}{}: {}".format(indent, k, v))


def prep_environment(args, cfg):
    """Output folder, logger and (optional) TensorBoard writer; logs the config and the arguments."""
    model_dir = '{}/{}'.format(cfg['model']['model_dir'], cfg['model']['name'])
    if args.rank == 0:
        prep_output_folder(model_dir, False)
    if getattr(args, 'distributed', False) and dist.is_initialized():
        dist.barrier()                     # the folder exists before any other rank opens its log
    logger = Logger(quiet=args.quiet, log_fn='{}/train.log'.format(model_dir), rank=args.rank)
    logger.add_line(str(datetime.datetime.now()))
    slurm = sorted(k for k in os.environ if 'SLURM' in k)
    if slurm:
        logger.add_line("=" * 30 + "   SLURM   " + "=" * 30)
        for k in slurm:
            logger.add_line('{:30}: {}'.format(k, os.environ[k]))
    logger.add_line("=" * 30 + "   Config   " + "=" * 30)
    _log_tree(logger, cfg)
    logger.add_line("=" * 30 + "   Args   " + "=" * 30)
    for k, v in vars(args).items():
        logger.add_line('{:30} {}'.format(k, v))
    tb_writter = None
    if cfg.get('log2tb') and args.rank == 0:
        try:
            from torch.utils.tensorboard import SummaryWriter
            tb_dir = '{}/tensorboard'.format(model_dir)
            os.makedirs(tb_dir, exist_ok=True)
            tb_writter = SummaryWriter(tb_dir)
        except Exception as e:            # tensorboard is an optional dependency here
            logger.add_line('TensorBoard disabled: {}'.format(e))
    return logger, tb_writter, model_dir


def build_model(cfg, logger=None):
    from .. import models
    assert cfg['arch'] in models.__dict__, 'Unknown model architecture'
    model = models.__dict__[cfg['arch']](**cfg['args'])
    if logger is not None:
        from .. import __version__, _lib
        logger.add_line("backend: avid_cma_b200 {} ({})".format(__version__, _lib.LIB_PATH))
        parts = model if isinstance(model, (list, tuple)) else [model]
        logger.add_line("=" * 30 + "   Model   " + "=" * 30)
        for m in parts:
            logger.add_line(str(m))
        logger.add_line("=" * 30 + "   Parameters   " + "=" * 30)
        for m in parts:
            logger.add_line(parameter_description(m))
    return model


def distribute_model_to_cuda(models, args, batch_size, num_workers, ngpus_per_node):
    """Placement of the model(s), chosen like main_utils.py:98-138: DistributedDataParallel when the job is distributed (pinned to
    args.gpu in the one-process-per-GPU case, in which the configured batch size / worker count are per NODE and get divided),
    plain .cuda(gpu) when a single GPU was requested, single-process DataParallel otherwise.  The towers' parameters are ordinary
    nn.Parameters whose .grad is filled by one autograd.Function per tower, so DDP's hooks and bucketed all-reduce work unchanged."""
    if ngpus_per_node == 0:
        return models, args, batch_size, num_workers
    pinned = args.gpu is not None
    if pinned:
        torch.cuda.set_device(args.gpu)

    def place(m):
        if args.distributed:
            if pinned and fused_grad_sync():
                # AVID_GRAD_SYNC=fused: gradients stay local, build_optimizer returns the optimizer that exchanges them over NVLink
                # peer memory inside its step (optim.ShardedAdam); the wrapper only provides `.module` and the `module.` key prefix
                from ..optim import LocalGradients
                return LocalGradients(m.cuda(args.gpu))
            ddp = torch.nn.parallel.DistributedDataParallel
            return ddp(m.cuda(args.gpu), device_ids=[args.gpu]) if pinned else ddp(m.cuda())
        if pinned:
            return m.cuda(args.gpu)
        # the reference wraps in DataParallel over ALL visible GPUs here; this package runs one process per GPU, so the wrap
        # (kept for `model.module`, main-avid.py:100) is pinned to the current device -- use --multiprocessing-distributed for more
        return torch.nn.DataParallel(m, device_ids=[torch.cuda.current_device()]).cuda()

    placed = [place(m) for m in models] if isinstance(models, list) else place(models)
    if args.distributed and pinned:
        batch_size, num_workers = int(batch_size / ngpus_per_node), int((num_workers + ngpus_per_node - 1) / ngpus_per_node)
    return placed, args, batch_size, num_workers


def fused_grad_sync():
    """AVID_GRAD_SYNC=fused (default: ddp, the reference's DistributedDataParallel + Adam): W > 1 ranks on one node exchange the
    gradients inside the optimizer step over NVLink peer memory (optim.ShardedAdam)."""
    if os.environ.get('AVID_GRAD_SYNC', 'ddp') != 'fused':
        return False
    from ..optim import ShardedAdam
    return ShardedAdam.available()


def build_dataloaders(cfg, num_workers, distributed, logger):
    train_loader = build_dataloader(cfg, cfg['train'], num_workers, distributed)
    logger.add_line("\n" + "=" * 30 + "   Train data   " + "=" * 30)
    logger.add_line(str(train_loader.dataset))
    return train_loader


def build_dataloader(db_cfg, split_cfg, num_workers, distributed):
    """DataLoader over {'frames', 'audio', 'index'} samples.  db_cfg['name'] == 'synthetic' builds SyntheticAV with the clip and
    spectrogram shapes the reference's transforms would produce from the same config keys (frames = clip duration x fps, crop
    size; spectrogram = audio duration x spectrogram_fps frames of n_fft / 2 + 1 bins, preprocessing.py:158-186)."""
    import torch.utils.data as data
    import torch.utils.data.distributed
    from ..datasets import SyntheticAV
    if db_cfg['name'] != 'synthetic':
        raise ValueError("dataset '{}': only 'synthetic' is built in (the PyAV / librosa datasets of the reference are out of scope; "
                         "feed their loader to run_phase instead)".format(db_cfg['name']))
    db = SyntheticAV(num_samples=db_cfg['num_samples'],
                     num_frames=int(db_cfg['video_clip_duration'] * db_cfg['video_fps']),
                     crop_size=db_cfg['crop_size'],
                     spectrogram=(int(db_cfg['audio_clip_duration'] * db_cfg['spectrogram_fps']), db_cfg['n_fft'] // 2 + 1),
                     clips_per_video=split_cfg.get('clips_per_video', 1), seed=db_cfg.get('seed', 0))
    sampler = torch.utils.data.distributed.DistributedSampler(db) if distributed else None
    return data.DataLoader(db, batch_size=db_cfg['batch_size'], shuffle=(sampler is None), drop_last=split_cfg['drop_last'],
                           num_workers=num_workers, pin_memory=torch.cuda.is_available(), sampler=sampler)


def build_criterion(cfg, logger=None):
    from .. import criterions
    criterion = criterions.__dict__[cfg['name']](**cfg['args'])
    if logger is not None:
        logger.add_line(str(criterion))
    return criterion


def build_optimizer(params, cfg, logger=None):
    """SGD / Adam + MultiStepLR from the reference's optimizer config (main_utils.py:240-259).  'adam' on CUDA parameters is the
    fused multi-tensor kernel (avid_cma_b200.optim.Adam: same update rule and state_dict layout as torch.optim.Adam)."""
    params, lr = list(params), cfg['lr']
    kind = cfg['name']
    if kind == 'sgd':
        optimizer = torch.optim.SGD(params, lr=lr['base_lr'], momentum=cfg['momentum'], weight_decay=cfg['weight_decay'], nesterov=cfg['nesterov'])
    elif kind == 'adam':
        from ..optim import Adam as FusedAdam, ShardedAdam
        on_gpu = bool(params) and all(p.is_cuda for p in params)
        cls = (ShardedAdam if fused_grad_sync() else FusedAdam) if on_gpu else torch.optim.Adam
        optimizer = cls(params, lr=lr['base_lr'], weight_decay=cfg['weight_decay'], betas=cfg.get('betas', [0.9, 0.999]))
    else:
        raise ValueError('Unknown optimizer.')
    return optimizer, torch.optim.lr_scheduler.MultiStepLR(optimizer, milestones=lr['milestones'], gamma=lr['gamma'])


def reference_state_dict(module):
    """state_dict in the reference's checkpoint layout.  A criterion whose memory banks are row-sharded over the ranks holds only
    its own rows; the full (N,128) banks are gathered (COLLECTIVE: every rank must call this).  Everything is moved to the CPU."""
    sd = module.state_dict()
    bank = getattr(module, 'nce_average', None)
    if bank is not None and getattr(bank, 'sharded', False):
        v, a = bank.full_banks()
        sd['nce_average.view1_mem'], sd['nce_average.view2_mem'] = v, a
    return {k: (t.detach().cpu() if torch.is_tensor(t) else t) for k, t in sd.items()}


class CheckpointManager(object):
    """{'epoch', 'model', 'optimizer', 'train_criterion'} files named like the reference's (main_utils.py:262-315): rank 0
    writes `checkpoint.pth.tar` (and copies the best one to `model_best.pth.tar`), every rank restores."""
    FILES = {'last': 'checkpoint.pth.tar', 'best': 'model_best.pth.tar'}

    def __init__(self, checkpoint_dir, rank=0):
        self.checkpoint_dir, self.rank, self.best_metric = checkpoint_dir, rank, 0.

    def _path(self, which):
        return '{}/{}'.format(self.checkpoint_dir, self.FILES[which])

    def last_checkpoint_fn(self):
        return self._path('last')

    def best_checkpoint_fn(self):
        return self._path('best')

    def checkpoint_fn(self, last=False, best=False):
        assert best != last, 'choose exactly one of last / best'
        return self._path('last' if last else 'best')

    def checkpoint_exists(self, last=False, best=False):
        return os.path.isfile(self.checkpoint_fn(last, best))

    def save(self, epoch, filename=None, eval_metric=0., **kwargs):
        """kwargs: name -> module / optimizer.  A criterion with row-sharded banks or an optimizer with sharded moments
        (optim.ShardedAdam) makes this call COLLECTIVE (the shards are gathered on every rank); otherwise ranks other than 0 return
        at once."""
        gather = any(getattr(getattr(m, 'nce_average', None), 'sharded', False) or getattr(m, 'collective_state_dict', False)
                     for m in kwargs.values())
        if self.rank != 0 and not gather:
            return
        state = {name: (reference_state_dict(m) if hasattr(m, 'nce_average') else m.state_dict()) for name, m in kwargs.items()}
        state['epoch'] = epoch
        if self.rank != 0:
            return
        improved = eval_metric > self.best_metric
        self.best_metric = max(self.best_metric, eval_metric)
        if filename is not None:
            save_checkpoint(state=state, is_best=False, filename='{}/{}'.format(self.checkpoint_dir, filename))
        else:
            save_checkpoint(state=state, is_best=improved, model_dir=self.checkpoint_dir)

    def restore(self, fn=None, restore_last=False, restore_best=False, **kwargs):
        path = fn if fn is not None else self.checkpoint_fn(restore_last, restore_best)
        ckp = torch.load(path, map_location='cpu', weights_only=False)      # reference checkpoints pickle plain dicts / numpy scalars
        for name, m in kwargs.items():
            m.load_state_dict(ckp[name], **({'strict': False} if name == 'train_criterion' else {}))
        return ckp['epoch']


def save_checkpoint(state, is_best, model_dir='.', filename=None):
    if filename is None:
        filename = '{}/checkpoint.pth.tar'.format(model_dir)
    torch.save(state, filename)
    if is_best:
        shutil.copyfile(filename, '{}/model_best.pth.tar'.format(model_dir))


def prep_output_folder(model_dir, evaluate):
    if evaluate:
        assert os.path.isdir(model_dir)
    else:
        os.makedirs(model_dir, exist_ok=True)


def parameter_description(model):
    rows = []
    for n, p in model.named_parameters():
        rows.append("{:70} | {:10} | {:30} | {}\n".format(n, 'Trainable' if p.requires_grad else 'Frozen', ' x '.join(str(s) for s in p.size()),
                                                          str(np.prod(p.size()))))
    return ''.join(rows)
