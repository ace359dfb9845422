#!/usr/bin/env python
"""AVID / AVID-CMA self-supervised training on the B200 path.

Command line, yaml config keys, log lines and checkpoint files are those of the reference's launcher (main-avid.py:24-46 flags;
:84-139 sequence: environment, model, DistributedDataParallel / DataParallel wrap, loader, criterion, optimizer + MultiStepLR,
optional resume, epochs with milestone and periodic checkpoints; :141-201 the training step), so existing configs and job
scripts keep working:

    python main_avid.py cfg.yaml [--quiet] [--seed S] [--gpu G]
    python main_avid.py cfg.yaml --multiprocessing-distributed --world-size 1 --rank 0 [--dist-url tcp://127.0.0.1:15475]
    torchrun --nproc-per-node N main_avid.py cfg.yaml --dist-url env://

`dataset.name: synthetic` (plus `dataset.num_samples`) trains on synthetic clips of the configured shapes; the other
differences from the reference's utils are listed in avid_cma_b200/utils/main_utils.py.
"""
import argparse
import os
import random
import time
import warnings

import torch
import torch.multiprocessing as mp
import yaml

from avid_cma_b200.utils import logger as logger_lib
from avid_cma_b200.utils import main_utils, metrics_utils

FLAGS = (
    (('cfg',), dict(help='yaml config')),
    (('--quiet',), dict(action='store_true')),
    (('--world-size',), dict(default=-1, type=int, help='number of nodes for distributed training')),
    (('--rank',), dict(default=-1, type=int, help='node rank for distributed training')),
    (('--dist-url',), dict(default='tcp://127.0.0.1:15475', type=str, help='url used to set up distributed training')),
    (('--dist-backend',), dict(default='nccl', type=str)),
    (('--seed',), dict(default=None, type=int)),
    (('--gpu',), dict(default=None, type=int, help='GPU id to use (disables data parallelism)')),
    (('--multiprocessing-distributed',), dict(action='store_true', help='one process per GPU of this node')),
)


def get_parser():
    ap = argparse.ArgumentParser(description='AVID / AVID-CMA training (B200 path)')
    for names, kwargs in FLAGS:
        ap.add_argument(*names, **kwargs)
    return ap


class Trainer:
    """Everything one worker process owns: model, loader, criterion, optimizer, scheduler, checkpoints."""

    def __init__(self, gpu, ngpus_per_node, args, cfg):
        args.gpu = gpu
        self.args = main_utils.initialize_distributed_backend(args, ngpus_per_node)
        self.cfg = cfg
        self.logger, self.tb_writter, self.model_dir = main_utils.prep_environment(self.args, cfg)
        data, loss = cfg['dataset'], cfg['loss']

        model = main_utils.build_model(cfg['model'], self.logger)
        self.model, self.args, data['batch_size'], cfg['num_workers'] = main_utils.distribute_model_to_cuda(
            model, self.args, data['batch_size'], cfg['num_workers'], ngpus_per_node)
        self.loader = main_utils.build_dataloaders(data, cfg['num_workers'], self.args.distributed, self.logger)

        self.device = self.args.gpu if self.args.gpu is not None else 0
        loss['args'].update(embedding_dim=getattr(self.model, 'module', self.model).out_dim, device=self.device)
        self.criterion = main_utils.build_criterion(loss, logger=self.logger)
        self.optimizer, self.scheduler = main_utils.build_optimizer(
            params=list(self.model.parameters()) + list(self.criterion.parameters()), cfg=cfg['optimizer'], logger=self.logger)
        self.checkpoints = main_utils.CheckpointManager(self.model_dir, rank=self.args.rank)
        self.first_epoch = self._maybe_resume()

    def _state(self):
        return dict(model=self.model, optimizer=self.optimizer, train_criterion=self.criterion)

    def _maybe_resume(self):
        if not self.cfg['resume']:
            return 0
        ck = self.checkpoints
        if not ck.checkpoint_exists(last=True):
            self.logger.add_line("No checkpoint found at '{}'".format(ck.last_checkpoint_fn()))
            return 0
        epoch = ck.restore(restore_last=True, **self._state())
        # the restored optimizer already carries the learning rate of `epoch`; only the scheduler's position moves
        self.scheduler.last_epoch = epoch
        self.scheduler._last_lr = [g['lr'] for g in self.optimizer.param_groups]
        self.logger.add_line("Checkpoint loaded: '{}' (epoch {})".format(ck.last_checkpoint_fn(), epoch))
        return epoch

    def fit(self):
        opt_cfg = self.cfg['optimizer']
        last_epoch, every = opt_cfg['num_epochs'], self.cfg.get('test_freq', 1)
        for epoch in range(self.first_epoch, last_epoch):
            if epoch in opt_cfg['lr']['milestones']:
                self.checkpoints.save(epoch, filename='checkpoint-ep{}.pth.tar'.format(epoch), **self._state())
            if self.args.distributed:
                self.loader.sampler.set_epoch(epoch)
            self.criterion.set_epoch(epoch)
            self.logger.add_line('=' * 30 + ' Epoch {} '.format(epoch) + '=' * 30)
            self.logger.add_line('LR: {}'.format(self.scheduler.get_last_lr()))
            run_phase('train', self.loader, self.model, self.optimizer, self.criterion, epoch, self.args, self.cfg, self.logger, self.tb_writter)
            self.scheduler.step()
            if epoch % every == 0 or epoch == last_epoch - 1:
                self.checkpoints.save(epoch + 1, **self._state())
        return self.model, self.criterion


def run_phase(phase, loader, model, optimizer, criterion, epoch, args, cfg, logger, tb_writter):
    """One pass over the loader (main-avid.py:141-201): H2D copy, both towers, criterion, loss.item(), backward, optimizer step,
    meters.  Returns the epoch's mean loss."""
    training = phase == 'train'
    meters = {'time': metrics_utils.AverageMeter('Time', ':6.3f', window_size=100),
              'data': metrics_utils.AverageMeter('Data', ':6.3f', window_size=100),
              'loss': metrics_utils.AverageMeter('Loss', ':.3e')}
    progress = logger_lib.ProgressMeter(len(loader), list(meters.values()), phase=phase, epoch=epoch, logger=logger, tb_writter=tb_writter)
    logger.add_line('\n{}: Epoch {}'.format(phase, epoch))
    model.train(training)
    device = args.gpu if args.gpu is not None else 0
    n_batches, tick = len(loader), time.time()
    for i, sample in enumerate(loader, start=1):
        meters['data'].update(time.time() - tick)
        video, audio, index = (sample[k].cuda(device, non_blocking=True) for k in ('frames', 'audio', 'index'))
        with torch.set_grad_enabled(training):
            video_emb, audio_emb = model(video, audio)
        loss, loss_debug = criterion(video_emb, audio_emb, index)
        meters['loss'].update(loss.item(), video.size(0))
        if training:
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
        meters['time'].update(time.time() - tick)
        tick = time.time()
        if i % cfg['print_freq'] == 0 or i == 1 or i == n_batches:
            progress.display(i)
            if tb_writter is not None:
                for key, val in loss_debug.items():
                    tb_writter.add_scalar('{}-batch/{}'.format(phase, key), float(val), epoch * n_batches + i - 1)
    if args.distributed:
        progress.synchronize_meters(args.gpu)
        progress.display(n_batches * args.world_size)
    if tb_writter is not None:
        for meter in progress.meters:
            tb_writter.add_scalar('{}-epoch/{}'.format(phase, meter.name), meter.avg, epoch)
    return meters['loss'].avg


def main_worker(gpu, ngpus_per_node, args, cfg):
    return Trainer(gpu, ngpus_per_node, args, cfg).fit()


def main(argv=None):
    args = get_parser().parse_args(argv)
    with open(args.cfg) as f:
        cfg = yaml.safe_load(f)
    if args.seed is not None:
        random.seed(args.seed)
        torch.manual_seed(args.seed)
        warnings.warn('Seeded run: host-side sampling is reproducible; kernels using atomics are not bit-reproducible.')
    if args.gpu is not None:
        warnings.warn('A specific GPU was chosen: data parallelism is disabled.')
    if args.dist_url == "env://" and args.world_size == -1:
        args.world_size = int(os.environ["WORLD_SIZE"])
    args.distributed = args.world_size > 1 or args.multiprocessing_distributed
    ngpus_per_node = torch.cuda.device_count()
    if args.multiprocessing_distributed:
        args.world_size *= ngpus_per_node                                    # one process per GPU of every node
        mp.spawn(main_worker, nprocs=ngpus_per_node, args=(ngpus_per_node, args, cfg))
        return None
    if args.dist_url == "env://" and args.gpu is None and "LOCAL_RANK" in os.environ:
        args.gpu = int(os.environ["LOCAL_RANK"])                             # torchrun: one process per GPU
    return main_worker(args.gpu, ngpus_per_node, args, cfg)


if __name__ == '__main__':
    main()
