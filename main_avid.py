#!/usr/bin/env python
"""AVID / AVID-CMA self-supervised training on the B200 path: the launcher of the reference (main-avid.py) re-stated on
avid_cma_b200.  Same command line, same yaml config keys, same sequence (main-avid.py:48-139: environment, model,
DistributedDataParallel / DataParallel wrap, loaders, criterion, optimizer + MultiStepLR, optional resume, epochs with
milestone / periodic checkpoints) and the same training step (run_phase, main-avid.py:141-201).

    python main_avid.py cfg.yaml [--quiet] [--seed S] [--gpu G]
    python main_avid.py cfg.yaml --multiprocessing-distributed --world-size 1 --rank 0 [--dist-url tcp://127.0.0.1:15475]
    torchrun --nproc-per-node N main_avid.py cfg.yaml --dist-url env://

dataset.name 'synthetic' (plus dataset.num_samples) trains on synthetic clips of the configured shapes; see
avid_cma_b200/utils/main_utils.py for the other differences from the reference's utils.
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


def get_parser():
    ap = argparse.ArgumentParser(description='AVID / AVID-CMA training (B200 path)')
    ap.add_argument('cfg', help='yaml config')
    ap.add_argument('--quiet', action='store_true')
    ap.add_argument('--world-size', default=-1, type=int, help='number of nodes for distributed training')
    ap.add_argument('--rank', default=-1, type=int, help='node rank for distributed training')
    ap.add_argument('--dist-url', default='tcp://127.0.0.1:15475', type=str, help='url used to set up distributed training')
    ap.add_argument('--dist-backend', default='nccl', type=str)
    ap.add_argument('--seed', default=None, type=int)
    ap.add_argument('--gpu', default=None, type=int, help='GPU id to use (disables data parallelism)')
    ap.add_argument('--multiprocessing-distributed', action='store_true', help='one process per GPU of this node')
    return ap


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
        args.world_size = ngpus_per_node * args.world_size
        mp.spawn(main_worker, nprocs=ngpus_per_node, args=(ngpus_per_node, args, cfg))
    else:
        if args.dist_url == "env://" and args.gpu is None and "LOCAL_RANK" in os.environ:
            args.gpu = int(os.environ["LOCAL_RANK"])        # torchrun: one process per GPU
        main_worker(args.gpu, ngpus_per_node, args, cfg)


def main_worker(gpu, ngpus_per_node, args, cfg):
    args.gpu = gpu
    args = main_utils.initialize_distributed_backend(args, ngpus_per_node)
    logger, tb_writter, model_dir = main_utils.prep_environment(args, cfg)

    model = main_utils.build_model(cfg['model'], logger)
    model, args, cfg['dataset']['batch_size'], cfg['num_workers'] = main_utils.distribute_model_to_cuda(
        model, args, cfg['dataset']['batch_size'], cfg['num_workers'], ngpus_per_node)
    train_loader = main_utils.build_dataloaders(cfg['dataset'], cfg['num_workers'], args.distributed, logger)

    device = args.gpu if args.gpu is not None else 0
    inner = model.module if hasattr(model, 'module') else model
    cfg['loss']['args']['embedding_dim'] = inner.out_dim
    cfg['loss']['args']['device'] = device
    train_criterion = main_utils.build_criterion(cfg['loss'], logger=logger)

    optimizer, scheduler = main_utils.build_optimizer(params=list(model.parameters()) + list(train_criterion.parameters()),
                                                      cfg=cfg['optimizer'], logger=logger)
    ckp_manager = main_utils.CheckpointManager(model_dir, rank=args.rank)

    start_epoch, end_epoch = 0, cfg['optimizer']['num_epochs']
    if cfg['resume']:
        if ckp_manager.checkpoint_exists(last=True):
            start_epoch = ckp_manager.restore(restore_last=True, model=model, optimizer=optimizer, train_criterion=train_criterion)
            # the restored optimizer already carries the learning rate of `start_epoch`; only the scheduler's position moves
            scheduler.last_epoch = start_epoch
            scheduler._last_lr = [g['lr'] for g in optimizer.param_groups]
            logger.add_line("Checkpoint loaded: '{}' (epoch {})".format(ckp_manager.last_checkpoint_fn(), start_epoch))
        else:
            logger.add_line("No checkpoint found at '{}'".format(ckp_manager.last_checkpoint_fn()))

    test_freq = cfg.get('test_freq', 1)
    for epoch in range(start_epoch, end_epoch):
        if epoch in cfg['optimizer']['lr']['milestones']:
            ckp_manager.save(epoch, model=model, train_criterion=train_criterion, optimizer=optimizer, filename='checkpoint-ep{}.pth.tar'.format(epoch))
        if args.distributed:
            train_loader.sampler.set_epoch(epoch)
        train_criterion.set_epoch(epoch)

        logger.add_line('=' * 30 + ' Epoch {} '.format(epoch) + '=' * 30)
        logger.add_line('LR: {}'.format(scheduler.get_last_lr()))
        run_phase('train', train_loader, model, optimizer, train_criterion, epoch, args, cfg, logger, tb_writter)
        scheduler.step()
        if epoch % test_freq == 0 or epoch == end_epoch - 1:
            ckp_manager.save(epoch + 1, model=model, optimizer=optimizer, train_criterion=train_criterion)
    return model, train_criterion


def run_phase(phase, loader, model, optimizer, criterion, epoch, args, cfg, logger, tb_writter):
    """One pass over the loader (main-avid.py:141-201): H2D copy, both towers, criterion, loss.item(), backward, optimizer."""
    logger.add_line('\n{}: Epoch {}'.format(phase, epoch))
    batch_time = metrics_utils.AverageMeter('Time', ':6.3f', window_size=100)
    data_time = metrics_utils.AverageMeter('Data', ':6.3f', window_size=100)
    loss_meter = metrics_utils.AverageMeter('Loss', ':.3e')
    progress = logger_lib.ProgressMeter(len(loader), [batch_time, data_time, loss_meter], phase=phase, epoch=epoch, logger=logger, tb_writter=tb_writter)
    training = phase == 'train'
    model.train(training)
    device = args.gpu if args.gpu is not None else 0
    end = time.time()
    for i, sample in enumerate(loader):
        data_time.update(time.time() - end)
        video = sample['frames'].cuda(device, non_blocking=True)
        audio = sample['audio'].cuda(device, non_blocking=True)
        index = sample['index'].cuda(device, non_blocking=True)
        with torch.set_grad_enabled(training):
            video_emb, audio_emb = model(video, audio)
        loss, loss_debug = criterion(video_emb, audio_emb, index)
        loss_meter.update(loss.item(), video.size(0))
        if training:
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
        batch_time.update(time.time() - end)
        end = time.time()
        if (i + 1) % cfg['print_freq'] == 0 or i == 0 or i + 1 == len(loader):
            progress.display(i + 1)
            if tb_writter is not None:
                step = epoch * len(loader) + i
                for key in loss_debug:
                    tb_writter.add_scalar('{}-batch/{}'.format(phase, key), float(loss_debug[key]), step)
    if args.distributed:
        progress.synchronize_meters(args.gpu)
        progress.display(len(loader) * args.world_size)
    if tb_writter is not None:
        for meter in progress.meters:
            tb_writter.add_scalar('{}-epoch/{}'.format(phase, meter.name), meter.avg, epoch)
    return loss_meter.avg


if __name__ == '__main__':
    main()
