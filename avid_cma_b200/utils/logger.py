"""Rank-0 line logger and progress display with the interface of the reference's utils/logger.py (Logger :15-41,
ProgressMeter :44-80).

One deliberate difference: the reference's ProgressMeter.synchronize_meters reads `self.progress.meters`, an attribute that
does not exist (logger.py:74,79), so an unchanged main-avid.py crashes at the end of the first distributed epoch
(main-avid.py:196).  Here `progress` is a property returning the meter itself, which makes that very line work."""
import datetime
import sys

import torch
from torch import distributed as dist


class Logger(object):
    def __init__(self, quiet=False, log_fn=None, rank=0, prefix=""):
        self.rank = 0 if rank is None else rank
        self.quiet, self.log_fn = quiet, log_fn
        self.prefix = prefix + ' | ' if prefix else ""
        self.file_pointers = []
        if self.rank == 0 and self.quiet:
            with open(log_fn, 'w'):
                pass

    def add_line(self, content):
        if self.rank != 0:
            return
        line = self.prefix + content
        if self.quiet:
            with open(self.log_fn, 'a') as f:
                f.write(line + '\n')
        else:
            print(line)
            sys.stdout.flush()


class ProgressMeter(object):
    def __init__(self, num_batches, meters, phase, epoch=None, logger=None, tb_writter=None):
        self.batches_per_epoch = num_batches
        self.meters, self.phase, self.epoch = meters, phase, epoch
        self.logger, self.tb_writter = logger, tb_writter
        width = len(str(num_batches // 1))
        head = '[{}]'.format(epoch) if epoch is not None else ''
        self.batch_fmtstr = head + '[{:' + str(width) + 'd}/' + ('{:' + str(width) + 'd}').format(num_batches) + ']'

    @property
    def progress(self):
        return self

    def display(self, batch):
        fields = ['{} | {} {}'.format(datetime.datetime.now(), self.phase, self.batch_fmtstr.format(batch))] + [str(m) for m in self.meters]
        line = '\t'.join(fields)
        if self.logger is None:
            print(line)
        else:
            self.logger.add_line(line)
        if self.tb_writter is not None:
            step = self.epoch * self.batches_per_epoch + batch
            for m in self.meters:
                self.tb_writter.add_scalar('{}-batch/{}'.format(self.phase, m.name), m.val, step)

    def synchronize_meters(self, cur_gpu):
        """Mean of every meter's average over the ranks (logger.py:73-80).  One all_gather_into_tensor of a small vector;
        works on the gloo backend too (CPU tensor when CUDA is not available)."""
        vals = torch.tensor([float(m.avg) for m in self.meters], dtype=torch.float32)
        if torch.cuda.is_available() and dist.get_backend() == 'nccl':
            vals = vals.cuda(cur_gpu)
        gathered = [torch.empty_like(vals) for _ in range(dist.get_world_size())]
        dist.all_gather(gathered, vals)
        mean = torch.stack(gathered).mean(0).cpu().tolist()
        for m, v in zip(self.meters, mean):
            m.avg = v
