"""AVID criterion on the fused CUDA NCE kernel (reference: criterions/avid.py).

Same constructor kwargs, forward signature, state_dict keys (`nce_average.view1_mem`,
`nce_average.view2_mem`, `criterion.avg_exp_score`) and loss semantics as the reference;
the ATen chain normalize -> gather -> bmm -> exp/log -> autograd is replaced by ONE pass of
`avid_nce_forward_backward` over the gathered bank rows (forward scores, loss terms and the
gradient w.r.t. the embeddings share the gather), and `update_memory` by `avid_bank_update`.
"""
import os
import pprint

import torch
from torch import nn
import torch.distributed as dist

from .. import ops
from ..utils.distributed_utils import _gather_from_all
from .nce import NCECriterion

__all__ = ['AVID']

# key name -> (ctx, bank): 0 = video embedding / view1_mem, 1 = audio embedding / view2_mem
_COMBO = {'v2a': (0, 1), 'a2v': (1, 0), 'v2v': (0, 0), 'a2a': (1, 1)}


class _FusedNCE(torch.autograd.Function):
    """loss_total, loss_keys = f(emb_v, emb_a); the kernel already produced d loss_total / d emb."""

    @staticmethod
    def forward(ctx, emb_v, emb_a, bank, y):
        loss_total, loss_keys, grad_v, grad_a = bank._run_fused(emb_v.detach().contiguous().float(),
                                                                emb_a.detach().contiguous().float(), y)
        ctx.save_for_backward(grad_v, grad_a)
        ctx.mark_non_differentiable(loss_keys)
        return loss_total, loss_keys

    @staticmethod
    def backward(ctx, g_total, _g_keys):
        grad_v, grad_a = ctx.saved_tensors
        return g_total * grad_v, g_total * grad_a, None, None


def _torch_device(device):
    return torch.device('cuda', device) if isinstance(device, int) else torch.device(device)


class AVIDSimilarityMemoryBank(nn.Module):
    """Memory banks + fused NCE.  Two distributed layouts:

    replicated (default, the reference's layout): every rank holds both (N,128) banks and update_memory
        all-gathers (embeddings, y) from all ranks (avid.py:103-129);
    sharded (AVID_SHARD_BANK=1 or shard=True): rank r holds rows [r*ceil(N/W), (r+1)*ceil(N/W)); per step the
        queries (embeddings, y) of all ranks are all-gathered once, every rank scores the positives / negatives
        it holds for all W*B queries, the partial gradients and loss terms are summed with one all-reduce, and
        each rank updates the rows it owns from the already gathered embeddings."""

    def __init__(self, memory_size, embedding_dim, xModal=True, wModal=False, num_negatives=1024, momentum=0.5, device=0, shard=None):
        super().__init__()
        if embedding_dim != 128:
            raise ValueError('the CUDA criterion kernels are specialised for embedding_dim == 128 (got %d)' % embedding_dim)
        self.num_negatives = num_negatives
        self.temperature = 0.07
        if not isinstance(momentum, (list, tuple)):
            momentum = [momentum] * 2
        self.momentum = momentum
        self.device = device
        self.memory_size = memory_size
        self.xModal = xModal
        self.wModal = wModal
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.distributed else 0
        self.world = dist.get_world_size() if self.distributed else 1
        if shard is None:
            shard = os.environ.get('AVID_SHARD_BANK', '0') == '1'
        self.sharded = bool(shard) and self.world > 1
        self.rows_per_rank = (memory_size + self.world - 1) // self.world if self.sharded else memory_size
        self.row_begin = min(memory_size, self.rank * self.rows_per_rank) if self.sharded else 0
        self.row_end = min(memory_size, self.row_begin + self.rows_per_rank)
        # counter-based sampler state: (seed, offset) of the Philox stream shared by ALL ranks' kernels.  The default generator's seed
        # differs from process to process unless the launcher seeds every worker, and a sharded step is only correct when every
        # rank draws the same negatives for a gathered query: rank 0's seed is broadcast.
        self._seed = self._shared_seed()
        self._offset = 0
        self._owner = None           # the AVID module (gives access to the NCECriterion and coefficients)
        self._ws = None
        self._gathered = None        # packed (emb_v, emb_a, y) of all ranks, kept from the sharded scoring pass for update_memory
        # raised by the NCE kernel when an instance index is outside [0, N) (the reference raises IndexError, avid.py:57-58); pinned
        # host memory, polled at the start of the next forward without a device sync
        self._bad = torch.zeros(1, dtype=torch.int32).pin_memory() if torch.cuda.is_available() else None
        self.init_memory(memory_size, embedding_dim)
        self._register_load_state_dict_pre_hook(self._on_load)
        self._register_state_dict_hook(self._on_save)

    def _shared_seed(self):
        seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        if self.distributed:
            dev = _torch_device(self.device) if dist.get_backend() == 'nccl' else torch.device('cpu')
            t = torch.tensor([seed - (1 << 64) if seed >= (1 << 63) else seed], dtype=torch.int64, device=dev)
            dist.broadcast(t, 0)
            seed = int(t.item()) & 0xFFFFFFFFFFFFFFFF
        return seed

    @staticmethod
    def _on_save(module, state_dict, prefix, local_metadata):
        """The Philox (seed, offset) travels with the checkpoint so that a resumed run does not replay the negative stream."""
        seed = module._seed - (1 << 64) if module._seed >= (1 << 63) else module._seed
        state_dict[prefix + 'sampler_state'] = torch.tensor([seed, module._offset], dtype=torch.int64)
        return state_dict

    # ---- reference API -------------------------------------------------------------------------
    def init_memory(self, num_items, embedding_dim):
        """avid.py:88-101: N(0,1) rows, L2-normalised, identical on every rank.  The reference draws them on rank 0 and broadcasts
        2 x (N,128); here row r is a function of (shared seed, bank, r) (avid_bank_init), so every rank -- replicated or owner of
        a shard -- fills its rows locally: no N x 128 tensor is ever materialised per shard and nothing crosses NVLink."""
        dev = _torch_device(self.device)
        rows = self.row_end - self.row_begin
        for which, name in enumerate(('view1_mem', 'view2_mem')):
            mem = torch.empty(max(rows, 0), embedding_dim, dtype=torch.float32, device=dev)
            if rows > 0:
                ops.bank_init_(mem, self.row_begin, self._seed, which)
            self.register_buffer(name, mem)
        if self.distributed:
            dist.barrier()

    def _on_load(self, state_dict, prefix, *args):
        """A checkpoint holds the full (N,128) banks (reference layout): a sharded rank loads its rows.  `sampler_state` (ours, absent
        from reference checkpoints) restores the Philox stream position."""
        st = state_dict.pop(prefix + 'sampler_state', None)
        if st is not None:
            self._seed, self._offset = int(st[0]) & 0xFFFFFFFFFFFFFFFF, int(st[1])
        if not self.sharded:
            return
        for name in ('view1_mem', 'view2_mem', 'positive_set'):
            key = prefix + name
            if name != 'positive_set' and key in state_dict and state_dict[key].shape[0] == self.memory_size:
                state_dict[key] = state_dict[key][self.row_begin:self.row_end]

    def full_banks(self):
        """(view1_mem, view2_mem) as full (N,128) tensors in the reference's checkpoint layout (all-gathers the shards)."""
        if not self.sharded:
            return self.view1_mem, self.view2_mem
        out = []
        for mem in (self.view1_mem, self.view2_mem):
            pad = torch.zeros(self.rows_per_rank, mem.shape[1], dtype=mem.dtype, device=mem.device)
            pad[:mem.shape[0]] = mem
            out.append(_gather_from_all(pad)[:self.memory_size])
        return tuple(out)

    def sample_negatives(self, y, K):
        """avid.py:82-86: (B,K) indices uniform over [0,N) minus {y_b}, drawn on the device (Philox4x32-10).

        Tests inject the reference's host-drawn indices by replacing this method on the instance."""
        B = y.shape[0]
        idx = ops.sample_negatives(y, K, self.memory_size, self._seed, self._offset + self.rank * B * K)
        self._offset += self.world * B * K
        return idx

    def _gather_queries(self, emb_v, emb_a, y, neg_idx=None):
        """ONE all-gather per step (the reference issues three, avid.py:109-111): every rank contributes one packed record
        [emb_v (B,128) f32 | emb_a (B,128) f32 | y (B) i64 | (injected negatives (B,K) i64)]; returns strided views
        (W,B,128), (W,B,128), (W,B)[, (W,B,K)] into the gathered buffer, rank-major like _gather_from_all."""
        B = emb_v.shape[0]
        nf, ny = B * 128, 2 * B
        nn_ = 2 * neg_idx.numel() if neg_idx is not None else 0
        rec = torch.empty(2 * nf + ny + nn_, dtype=torch.float32, device=emb_v.device)
        rec[:nf].copy_(emb_v.reshape(-1))
        rec[nf:2 * nf].copy_(emb_a.reshape(-1))
        rec[2 * nf:2 * nf + ny].view(torch.int64).copy_(y)
        if neg_idx is not None:
            rec[2 * nf + ny:].view(torch.int64).copy_(neg_idx.reshape(-1))
        W = self.world
        flat = torch.empty(W * rec.numel(), dtype=torch.float32, device=emb_v.device)
        dist.all_gather_into_tensor(flat, rec)
        allrec = flat.view(W, rec.numel())
        out = [allrec[:, :nf].view(W, B, 128), allrec[:, nf:2 * nf].view(W, B, 128), allrec[:, 2 * nf:2 * nf + ny].view(torch.int64)]
        if neg_idx is not None:
            out.append(allrec[:, 2 * nf + ny:].view(torch.int64).view(W, B, neg_idx.shape[1]))
        return out

    def update_memory(self, video_emb, audio_emb, y):
        """avid.py:103-129.  Takes the un-normalised embeddings: the kernel normalises them itself."""
        if self._gathered is not None:
            video_emb, audio_emb, y = self._gathered
            self._gathered = None
        elif self.distributed:
            video_emb, audio_emb, y = self._gather_queries(video_emb, audio_emb, y)
        ops.bank_update(self.view1_mem, self.view2_mem, video_emb, audio_emb, y, float(self.momentum[0]), float(self.momentum[1]),
                        row_begin=self.row_begin, row_end=self.row_end)

    def _check_indices(self):
        if self._bad is not None and int(self._bad[0]) != 0:
            self._bad.zero_()
            raise IndexError('an instance index passed to the criterion was outside [0, %d)' % self.memory_size)

    # ---- fused path ----------------------------------------------------------------------------
    def keys(self):
        """[(name, ctx, bank, pos_mode, num_neg)] in the insertion order of avid.py:69-75."""
        K = int(self.num_negatives)
        out = []
        if self.xModal:
            out += [('v2a',) + _COMBO['v2a'] + (0, K), ('a2v',) + _COMBO['a2v'] + (0, K)]
        if self.wModal:
            out += [('v2v',) + _COMBO['v2v'] + (0, K), ('a2a',) + _COMBO['a2a'] + (0, K)]
        return out

    def _positive_set(self):
        return None

    def _sampler_overridden(self):
        return 'sample_negatives' in self.__dict__

    def _draw(self, y):
        """Returns (neg_idx or None, seed, offset): None means 'draw inside the kernel'.  The Philox counter of query b of
        rank r in step s is offset_s + (r*B + b)*K + k, i.e. the gathered batch of a sharded step draws exactly the
        negatives the ranks of a replicated step draw."""
        K = int(self.num_negatives)
        if self._sampler_overridden():
            return self.sample_negatives(y, K).to(device=y.device, dtype=torch.int64).contiguous(), 0, 0
        off = self._offset
        self._offset += self.world * y.shape[0] * K
        return None, self._seed, off

    def _key_tuples(self):
        keys = self.keys()
        weights = self._owner._key_weights([k[0] for k in keys])
        return keys, [(k[1], k[2], k[3], k[4], w) for k, w in zip(keys, weights)]

    def _workspace(self, B, K, pos_k, nkeys, device):
        if self._ws is None or self._ws_shape != (B, K, pos_k, nkeys):
            self._ws = ops.nce_workspace(B, K, pos_k, nkeys, device)
            self._ws_shape = (B, K, pos_k, nkeys)
        return self._ws

    def _run_fused(self, emb_v, emb_a, y):
        y = y.to(device=emb_v.device, dtype=torch.int64).contiguous()
        if self.sharded:
            return self._run_sharded(emb_v, emb_a, y)
        crit = self._owner.criterion
        keys, key_tuples = self._key_tuples()
        B, K = emb_v.shape[0], int(self.num_negatives)
        pos = self._positive_set()
        pos_k = pos.shape[1] if pos is not None else 0
        ws = self._workspace(B, K, pos_k, len(keys), emb_v.device)
        neg_idx, seed, offset = self._draw(y)
        if neg_idx is None:
            offset += self.rank * B * K
        out = torch.empty(1 + len(keys) + 2 * B * 128, dtype=torch.float32, device=emb_v.device)
        loss_total, loss_keys = out[0:1], out[1:1 + len(keys)]
        grad_v = out[1 + len(keys):1 + len(keys) + B * 128].view(B, 128)
        grad_a = out[1 + len(keys) + B * 128:].view(B, 128)
        args = ops.make_nce_args(emb_v, emb_a, y, self.view1_mem, self.view2_mem, key_tuples, K, crit.avg_exp_score,
                                 neg_idx=neg_idx, seed=seed, offset=offset, positive_set=pos, temperature=self.temperature,
                                 loss_keys=loss_keys, loss_total=loss_total, grad_v=grad_v, grad_a=grad_a, bad_index=self._bad)
        if not crit.z_ready():
            # nce.py:21-36: Z <- mean exp(score) over the negatives of the first key of the first batch
            # (mean of the per-rank means when distributed), frozen afterwards.
            z = torch.empty(1, dtype=torch.float32, device=emb_v.device)
            ops.nce_partition_mean(args, 0, z, ws)
            if self.distributed:
                dist.all_reduce(z)
                z /= dist.get_world_size()
            crit.avg_exp_score.copy_(z.reshape(()))
            crit.mark_ready()
        ops.nce_forward_backward(args, ws)
        return loss_total.reshape(()), loss_keys, grad_v, grad_a

    def _reduce_scatter(self, mine, part):
        """mine (rec) <- sum over ranks of part[rank] (W, rec)."""
        if dist.get_backend() == 'gloo':        # gloo (CPU protocol tests) has no reduce-scatter
            dist.all_reduce(part)
            mine.copy_(part[self.rank])
        else:
            dist.reduce_scatter_tensor(mine, part.view(-1))

    def _run_sharded(self, emb_v, emb_a, y):
        """One step of the row-partitioned protocol: ONE packed all-gather of the queries -> score owned rows
        -> ONE reduce-scatter of the partials -> finalize own queries.  Scores / gradients cross NVLink (4 B per negative),
        bank rows (512 B) never do."""
        crit = self._owner.criterion
        keys, key_tuples = self._key_tuples()
        W, B, K, nk = self.world, emb_v.shape[0], int(self.num_negatives), len(keys)
        dev = emb_v.device
        pos = self._positive_set()
        pos_k = pos.shape[1] if pos is not None else 0
        neg_idx, seed, offset = self._draw(y)
        # (1) queries of all ranks, rank-major (these are also what update_memory needs: avid.py:109-111)
        gathered = self._gather_queries(emb_v, emb_a, y, neg_idx)
        all_v, all_a, all_y = gathered[:3]
        all_neg = gathered[3] if neg_idx is not None else None
        self._gathered = (all_v, all_a, all_y)
        ws = self._workspace(W * B, K, pos_k, nk, dev)
        # (2) partial dL/d(normalised embedding) and per-query loss terms over the rows this rank holds, written straight into the
        #     rank-major records the reduce-scatter sends back to the query owners
        nf = B * 128
        part = torch.empty(W, 2 * nf + nk * B, dtype=torch.float32, device=dev)
        gh_v, gh_a = part[:, :nf].view(W, B, 128), part[:, nf:2 * nf].view(W, B, 128)
        loss_part = part[:, 2 * nf:].view(W, nk, B)
        args = ops.make_nce_args(all_v, all_a, all_y, self.view1_mem, self.view2_mem, key_tuples, K, crit.avg_exp_score,
                                 num_rows=self.memory_size, row_begin=self.row_begin, row_end=self.row_end, neg_idx=all_neg, seed=seed,
                                 offset=offset, positive_set=pos, mean_batch=B, temperature=self.temperature,
                                 grad_hat_v=gh_v, grad_hat_a=gh_a, loss_part=loss_part, bad_index=self._bad)
        if not crit.z_ready():
            # the sharded partition pass returns the SUM of exp(score) over held rows; equal batches make the mean over all
            # W*B*K negatives equal to the reference's mean of per-rank means (nce.py:27-33)
            z = torch.empty(1, dtype=torch.float32, device=dev)
            ops.nce_partition_mean(args, 0, z, ws)
            dist.all_reduce(z)
            z /= float(W * B * keys[0][4])
            crit.avg_exp_score.copy_(z.reshape(()))
            crit.mark_ready()
        ops.nce_forward_backward(args, ws)
        # (3) sum the partials over the shards; every rank receives the record of its own B queries
        mine = torch.empty(2 * nf + nk * B, dtype=torch.float32, device=dev)
        self._reduce_scatter(mine, part)
        # (4) backward of F.normalize, batch means and coefficient mix for this rank's own queries
        out = torch.empty(1 + nk + 2 * B * 128, dtype=torch.float32, device=dev)
        loss_total, loss_keys = out[0:1], out[1:1 + nk]
        grad_v, grad_a = out[1 + nk:1 + nk + B * 128].view(B, 128), out[1 + nk + B * 128:].view(B, 128)
        fin = ops.make_nce_args(emb_v, emb_a, y, self.view1_mem, self.view2_mem, key_tuples, K, crit.avg_exp_score,
                                num_rows=self.memory_size, row_begin=self.row_begin, row_end=self.row_end, mean_batch=B,
                                temperature=self.temperature, loss_keys=loss_keys, loss_total=loss_total, grad_v=grad_v, grad_a=grad_a,
                                grad_hat_v=mine[:nf].view(B, 128), grad_hat_a=mine[nf:2 * nf].view(B, 128), loss_part=mine[2 * nf:].view(nk, B))
        ops.nce_finalize(fin, ws)
        return loss_total.reshape(()), loss_keys, grad_v, grad_a

    def forward(self, video_emb, audio_emb, y):
        """Returns (loss_total, {key: loss}) -- the reference returns raw scores here (avid.py:47-80) and leaves
        the loss to NCECriterion; the fused kernel produces both at once, then the bank is updated (avid.py:78)."""
        self._check_indices()
        loss_total, loss_keys = _FusedNCE.apply(video_emb, audio_emb, self, y)
        with torch.no_grad():
            self.update_memory(video_emb.detach().contiguous().float(), audio_emb.detach().contiguous().float(),
                               y.to(device=video_emb.device, dtype=torch.int64).contiguous())
        return loss_total, {k[0]: loss_keys[i] for i, k in enumerate(self.keys())}

    def __repr__(self):
        repr_dict = {
            'name': self._get_name(),
            'num_negatives': int(self.num_negatives),
            'momentum': [float(self.momentum[0]), float(self.momentum[1])],
            'view1_buffer_size': self.view1_mem.shape,
            'view2_buffer_size': self.view2_mem.shape,
        }
        return pprint.pformat(repr_dict, indent=2)


def _restore_bank_and_partition(module, checkpoint):
    """avid.py:187-200 / avid_cma.py:308-319: warm-start the banks and Z from another run's checkpoint."""
    ckp = torch.load(checkpoint, map_location='cpu', weights_only=False)['train_criterion']
    state_dict = module.state_dict()
    state_dict['nce_average.view1_mem'] = ckp['nce_average.view1_mem']
    state_dict['nce_average.view2_mem'] = ckp['nce_average.view2_mem']
    Z = torch.stack([ckp[k].reshape(()).float() for k in ckp if 'avg_exp_score' in k]).mean()
    for k in state_dict:
        if 'avg_exp_score' in k:
            state_dict[k] = Z
    module.load_state_dict(state_dict)


class AVID(nn.Module):
    def __init__(self, num_data, embedding_dim, num_negatives=4096, momentum=0.9, xModal_coeff=1., wModal_coeff=0.,
                 checkpoint=None, device=0):
        super().__init__()
        self.nce_average = AVIDSimilarityMemoryBank(memory_size=num_data, embedding_dim=embedding_dim, num_negatives=num_negatives,
                                                    momentum=momentum, xModal=xModal_coeff > 0., wModal=wModal_coeff > 0., device=device)
        self.nce_average = self.nce_average.to(_torch_device(device))
        object.__setattr__(self.nce_average, '_owner', self)
        sum_coeff = xModal_coeff + wModal_coeff
        self.xModal_coeff = xModal_coeff / sum_coeff
        self.wModal_coeff = wModal_coeff / sum_coeff
        self.criterion = NCECriterion(num_data).to(_torch_device(device))
        if checkpoint is not None:
            _restore_bank_and_partition(self, checkpoint)

    def _key_weights(self, names):
        # avid.py:216-233: each pair of directions is averaged, then mixed with the normalised coefficients
        return [(self.xModal_coeff if n in ('v2a', 'a2v') else self.wModal_coeff) / 2. for n in names]

    def forward(self, emb1, emb2, target):
        """emb1: video embeddings (N, D); emb2: audio embeddings (N, D); target: instance labels (N)."""
        total_loss, losses = self.nce_average(emb1, emb2, target)
        tb_log = {}
        zero = total_loss.new_zeros(())
        xModal_loss, wModal_loss = zero, zero
        for k, loss in losses.items():
            if k in {'v2a', 'a2v'}:
                xModal_loss = xModal_loss + loss / 2.
            elif k in {'v2v', 'a2a'}:
                wModal_loss = wModal_loss + loss / 2.
            tb_log[f'Loss/{k}'] = loss
        tb_log['Loss/xModal'] = xModal_loss
        tb_log['Loss/wModal'] = wModal_loss
        return total_loss, tb_log

    def set_epoch(self, epoch):
        pass
