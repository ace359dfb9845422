"""AVID-CMA criterion on the fused CUDA kernels (reference: criterions/avid_cma.py).

Positive mining (CMASampler, avid_cma.py:24-123: one worker process per GPU fed through
multiprocessing queues, the whole bank re-streamed per 16 queries) becomes an in-process tiled
similarity + streaming top-k kernel (csrc/cma.cu); in distributed runs every rank mines a slice of
the queries and the slices are all-gathered (instead of rank 0 mining everything and broadcasting).
The positive-aware negative remap (avid_cma.py:196-209) happens inside the sampler / NCE kernel.
"""
import torch
from torch import nn
import torch.distributed as dist

from .. import ops
from .avid import AVIDSimilarityMemoryBank, _COMBO, _restore_bank_and_partition, _torch_device
from .nce import NCECriterion

__all__ = ['AVID_CMA']


class AVIDSimilarityPositiveExpansion(AVIDSimilarityMemoryBank):
    def __init__(self, memory_size, embedding_dim, xModalInst=True, wModalInst=False, xModalPos=False, wModalPos=True,
                 num_negatives=1024, num_negatives_within=None, sampling_args=None, momentum=0.5, device=0):
        super().__init__(memory_size=memory_size, embedding_dim=embedding_dim, xModal=xModalInst, wModal=wModalInst,
                         num_negatives=num_negatives, momentum=momentum, device=device)
        self.num_negatives_within = num_negatives_within
        self.sampling_args = sampling_args
        self.xModalInst = xModalInst
        self.wModalInst = wModalInst
        self.xModalPos = xModalPos
        self.wModalPos = wModalPos

    def keys(self):
        """Score keys in the insertion order of avid_cma.py:169-188 (including the reference's quirk that the
        wModalInst branch re-writes the cross-modal 'inst-v2a'/'inst-a2v' keys, avid_cma.py:175-177)."""
        K = int(self.num_negatives)
        Kw = K if self.num_negatives_within is None else int(self.num_negatives_within)
        out = []
        if self.xModalInst or self.wModalInst:
            out += [('inst-v2a',) + _COMBO['v2a'] + (0, K), ('inst-a2v',) + _COMBO['a2v'] + (0, K)]
        if self.xModalPos:
            out += [('pos-v2a',) + _COMBO['v2a'] + (1, K), ('pos-a2v',) + _COMBO['a2v'] + (1, K)]
        if self.wModalPos:
            out += [('pos-v2v',) + _COMBO['v2v'] + (1, Kw), ('pos-a2a',) + _COMBO['a2a'] + (1, Kw)]
        return out

    def _positive_set(self):
        return getattr(self, 'positive_set', None)

    def _sampler_overridden(self):
        return 'memory_sampling' in self.__dict__

    def _draw(self, y):
        if self._sampler_overridden():
            _, neg = self.memory_sampling(y)
            return neg.to(device=y.device, dtype=torch.int64).contiguous(), 0, 0
        off = self._offset
        self._offset += self.world * y.shape[0] * int(self.num_negatives)
        return None, self._seed, off

    def memory_sampling(self, y):
        """avid_cma.py:196-209: (positive indices (B,pos_k) int64, negative indices (B,K) int64 that avoid them)."""
        B, K = y.shape[0], int(self.num_negatives)
        pos = self.positive_set[y].long()
        neg = ops.sample_negatives(y, K, self.memory_size, self._seed, self._offset + self.rank * B * K, self.positive_set)
        self._offset += self.world * B * K
        return pos, neg

    def _candidate_shards(self):
        """(video rows, audio rows, first row) of every bank shard in turn; a sharded run streams the other ranks' rows
        through one broadcast buffer, so no rank ever holds more than its own shard plus one visiting shard."""
        if not self.sharded:
            yield self.view1_mem, self.view2_mem, 0
            return
        N, per = self.memory_size, self.rows_per_rank
        for r in range(self.world):
            lo, hi = min(N, r * per), min(N, (r + 1) * per)
            if hi <= lo:
                continue
            if r == self.rank:
                cv, ca = self.view1_mem, self.view2_mem
            else:
                cv = torch.empty(hi - lo, 128, dtype=torch.float32, device=self.view1_mem.device)
                ca = torch.empty_like(cv)
            dist.broadcast(cv, r)
            dist.broadcast(ca, r)
            yield cv, ca, lo

    def _any_rank(self, flag):
        """OR of a boolean over the ranks (the shard stream of a sharded bank is collective: every rank must walk it again
        if any rank has queries to re-mine)."""
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=self.view1_mem.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(int(t.item()))

    def find_correspondences(self):
        """avid_cma.py:211-229: every rank mines the positives of N/W queries against all N candidates; the slices are
        all-gathered (the reference mines everything on rank 0 and broadcasts)."""
        pos_k = self.sampling_args['pos_k']
        if pos_k <= 0:
            return
        N = self.memory_size
        world = self.world
        per = (N + world - 1) // world
        lo, hi = min(N, self.rank * per), min(N, (self.rank + 1) * per)
        dev = self.view1_mem.device
        mine = torch.zeros(per, pos_k, dtype=torch.int32, device=dev)
        if self.sharded:
            qv, qa = self.view1_mem, self.view2_mem          # the rows this rank owns are its queries
        else:
            qv, qa = self.view1_mem[lo:hi], self.view2_mem[lo:hi]
        if hi > lo or self.sharded:
            res = ops.cma_topk(qv, qa, self._candidate_shards, pos_k, self.sampling_args['type'],
                               any_rank=self._any_rank if self.sharded else None)
            mine[:hi - lo] = res
        if self.distributed:
            full = torch.empty(world * per, pos_k, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(full, mine)
            positive_set = full[:N].contiguous()
        else:
            positive_set = mine[:N]
        self.register_buffer('positive_set', positive_set)
        if self.distributed:
            dist.barrier()


class AVID_CMA(nn.Module):
    def __init__(self, num_data, embedding_dim, num_negatives=1024, num_negatives_within=None, momentum=0.5,
                 xModalInstCoeff=1., wModalInstCoeff=0., xModalPosCoeff=0., wModalPosCoeff=1., sampling_args=None,
                 checkpoint=None, resample_freq=-1, device=0):
        super().__init__()
        self.nce_average = AVIDSimilarityPositiveExpansion(
            memory_size=num_data, embedding_dim=embedding_dim, num_negatives=num_negatives, num_negatives_within=num_negatives_within,
            momentum=momentum, xModalInst=xModalInstCoeff > 0., xModalPos=xModalPosCoeff > 0., wModalInst=wModalInstCoeff > 0.,
            wModalPos=wModalPosCoeff > 0., sampling_args=sampling_args, device=device)
        self.nce_average = self.nce_average.to(_torch_device(device))
        object.__setattr__(self.nce_average, '_owner', self)
        sum_coeff = xModalInstCoeff + wModalInstCoeff + xModalPosCoeff + wModalPosCoeff
        self.xModalInstCoeff = xModalInstCoeff / sum_coeff
        self.wModalInstCoeff = wModalInstCoeff / sum_coeff
        self.xModalPosCoeff = xModalPosCoeff / sum_coeff
        self.wModalPosCoeff = wModalPosCoeff / sum_coeff
        self.criterion = NCECriterion(num_data).to(_torch_device(device))
        if checkpoint is not None:
            _restore_bank_and_partition(self, checkpoint)
        self.resample_freq = resample_freq
        self.nce_average.find_correspondences()

    def _key_weights(self, names):
        # avid_cma.py:338-359
        w = {'inst-v2a': self.xModalInstCoeff, 'inst-a2v': self.xModalInstCoeff, 'inst-v2v': self.wModalInstCoeff,
             'inst-a2a': self.wModalInstCoeff, 'pos-v2a': self.xModalPosCoeff, 'pos-a2v': self.xModalPosCoeff,
             'pos-v2v': self.wModalPosCoeff, 'pos-a2a': self.wModalPosCoeff}
        return [w[n] / 2. for n in names]

    def forward(self, emb1, emb2, target):
        total_loss, losses = self.nce_average(emb1, emb2, target)
        tb_log = {f'Loss/{k}': v for k, v in losses.items()}
        return total_loss, tb_log

    def set_epoch(self, epoch):
        # avid_cma.py:361-364
        if self.resample_freq > 0 and epoch > 0 and epoch % self.resample_freq == 0:
            self.nce_average.find_correspondences()
