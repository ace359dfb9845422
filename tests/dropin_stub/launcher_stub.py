"""TEST FIXTURE: the call sequence of the reference launcher (main-avid.py:48-201), abbreviated, importing everything under the
REFERENCE's module names (`utils.logger`, `utils.main_utils`, `utils.metrics_utils`) -- /root/reference does not exist on the GPU
box, so this stands in for `python main-avid.py ...` when the drop-in redirect (avid_cma_b200/dropin_site/sitecustomize.py) is
tested there.  Nothing in here names avid_cma_b200."""
import argparse
import time

import torch
import torch.multiprocessing as mp
import yaml

import utils.logger                      # main-avid.py:20
from utils import main_utils             # main-avid.py:21


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('cfg')
    ap.add_argument('--quiet', action='store_true')
    ap.add_argument('--world-size', default=-1, type=int)
    ap.add_argument('--rank', default=-1, type=int)
    ap.add_argument('--dist-url', default='tcp://localhost:15475')
    ap.add_argument('--dist-backend', default='nccl')
    ap.add_argument('--seed', default=None, type=int)
    ap.add_argument('--gpu', default=None, type=int)
    ap.add_argument('--multiprocessing-distributed', action='store_true')
    return ap.parse_args()


def worker(gpu, ngpus, args, cfg):                                   # main-avid.py:84-138
    args.gpu = gpu
    args = main_utils.initialize_distributed_backend(args, ngpus)
    logger, tb, model_dir = main_utils.prep_environment(args, cfg)
    model = main_utils.build_model(cfg['model'], logger)
    model, args, cfg['dataset']['batch_size'], cfg['num_workers'] = main_utils.distribute_model_to_cuda(
        model, args, cfg['dataset']['batch_size'], cfg['num_workers'], ngpus)
    loader = main_utils.build_dataloaders(cfg['dataset'], cfg['num_workers'], args.distributed, logger)
    device = args.gpu if args.gpu is not None else 0
    cfg['loss']['args']['embedding_dim'] = model.module.out_dim      # main-avid.py:100 (needs a wrapped model)
    cfg['loss']['args']['device'] = device
    criterion = main_utils.build_criterion(cfg['loss'], logger=logger)
    optimizer, scheduler = main_utils.build_optimizer(params=list(model.parameters()) + list(criterion.parameters()),
                                                      cfg=cfg['optimizer'], logger=logger)
    ckp = main_utils.CheckpointManager(model_dir, rank=args.rank)
    for epoch in range(cfg['optimizer']['num_epochs']):
        if args.distributed:
            loader.sampler.set_epoch(epoch)
        scheduler.step(epoch)
        criterion.set_epoch(epoch)
        logger.add_line('=' * 30 + ' Epoch {} '.format(epoch) + '=' * 30)
        phase(loader, model, optimizer, criterion, epoch, args, cfg, logger, tb)
        ckp.save(epoch + 1, model=model, optimizer=optimizer, train_criterion=criterion)


def phase(loader, model, optimizer, criterion, epoch, args, cfg, logger, tb):      # main-avid.py:141-201
    from utils import metrics_utils
    meters = [metrics_utils.AverageMeter('Time', ':6.3f', window_size=100), metrics_utils.AverageMeter('Loss', ':.3e')]
    progress = utils.logger.ProgressMeter(len(loader), meters, phase='train', epoch=epoch, logger=logger, tb_writter=tb)
    model.train(True)
    device = args.gpu if args.gpu is not None else 0
    end = time.time()
    for i, sample in enumerate(loader):
        video = sample['frames'].cuda(device, non_blocking=True)
        audio = sample['audio'].cuda(device, non_blocking=True)
        index = sample['index'].cuda(device, non_blocking=True)
        video_emb, audio_emb = model(video, audio)
        loss, loss_debug = criterion(video_emb, audio_emb, index)
        meters[1].update(loss.item(), video.size(0))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        meters[0].update(time.time() - end)
        end = time.time()
        if (i + 1) % cfg['print_freq'] == 0 or i == 0 or i + 1 == len(loader):
            progress.display(i + 1)
            assert all(hasattr(v, 'item') for v in loss_debug.values())
    if args.distributed:
        progress.synchronize_meters(args.gpu)                        # crashes in the reference (logger.py:74): must work here
        progress.display(len(loader) * args.world_size)


if __name__ == '__main__':
    a = parse()
    config = yaml.safe_load(open(a.cfg))
    if a.seed is not None:
        torch.manual_seed(a.seed)
    a.distributed = a.world_size > 1 or a.multiprocessing_distributed
    n = torch.cuda.device_count()
    if a.multiprocessing_distributed:
        a.world_size = n * a.world_size
        mp.spawn(worker, nprocs=n, args=(n, a, config))              # main-avid.py:78: one process per GPU
    else:
        worker(a.gpu, n, a, config)
