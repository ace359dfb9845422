"""Oracle for the criterion half of the hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates, on CPU tensors of any float dtype (fp32 to mirror the reference, fp64 to
budget rounding):

  * NCECriterion.forward / compute_partition_function      criterions/nce.py:21-58
  * AVIDSimilarityMemoryBank.forward / sample_negatives /
    update_memory                                          criterions/avid.py:47-129
  * AVID.forward coefficient mix                           criterions/avid.py:202-233
  * AVIDSimilarityPositiveExpansion.forward /
    memory_sampling                                        criterions/avid_cma.py:150-209
  * AVID_CMA.forward coefficient mix                       criterions/avid_cma.py:325-359
  * CMASampler.sample_instance                             criterions/avid_cma.py:42-73
"""
from collections import namedtuple, OrderedDict

import numpy as np
import torch

TEMPERATURE = 0.07  # avid.py:32

# ctx / bank: 'v' (video embedding, view1_mem) or 'a' (audio embedding, view2_mem)
# pos: 'self' (row y) or 'set' (rows positive_set[y]); num_neg: first num_neg shared negatives
Key = namedtuple("Key", "name ctx bank pos num_neg weight")


def l2_normalize(x, eps=1e-12):
    """F.normalize(x, p=2, dim=1): x / max(||x||, eps)  (avid.py:52-53,92,95,122,128)."""
    n = x.pow(2).sum(1, keepdim=True).sqrt().clamp_min(eps)
    return x / n


def avid_keys(num_negatives, xModal_coeff=1.0, wModal_coeff=0.0):
    """Score keys of AVID in the insertion order of avid.py:69-75 with the weights of avid.py:181-183,216-233."""
    s = xModal_coeff + wModal_coeff
    xw, ww = xModal_coeff / s, wModal_coeff / s
    keys = []
    if xModal_coeff > 0:
        keys += [Key("v2a", "v", "a", "self", num_negatives, xw / 2), Key("a2v", "a", "v", "self", num_negatives, xw / 2)]
    if wModal_coeff > 0:
        keys += [Key("v2v", "v", "v", "self", num_negatives, ww / 2), Key("a2a", "a", "a", "self", num_negatives, ww / 2)]
    return keys


def avid_cma_keys(num_negatives, num_negatives_within=None, xModalInstCoeff=1.0, wModalInstCoeff=0.0,
                  xModalPosCoeff=0.0, wModalPosCoeff=1.0):
    """Score keys of AVID_CMA (avid_cma.py:169-188) and their weights (avid_cma.py:297-301,338-359).

    Quirk kept from the reference: the wModalInst branch re-writes the 'inst-v2a'/'inst-a2v'
    keys with cross-modal scores (avid_cma.py:175-177), and the loss loop then files them under
    the xModalInst coefficient; 'inst-v2v'/'inst-a2a' never exist.
    """
    s = xModalInstCoeff + wModalInstCoeff + xModalPosCoeff + wModalPosCoeff
    xi, xp, wp = xModalInstCoeff / s, xModalPosCoeff / s, wModalPosCoeff / s
    kw = num_negatives if num_negatives_within is None else num_negatives_within
    keys = []
    if xModalInstCoeff > 0 or wModalInstCoeff > 0:
        keys += [Key("inst-v2a", "v", "a", "self", num_negatives, xi / 2),
                 Key("inst-a2v", "a", "v", "self", num_negatives, xi / 2)]
    if xModalPosCoeff > 0:
        keys += [Key("pos-v2a", "v", "a", "set", num_negatives, xp / 2),
                 Key("pos-a2v", "a", "v", "set", num_negatives, xp / 2)]
    if wModalPosCoeff > 0:
        keys += [Key("pos-v2v", "v", "v", "set", kw, wp / 2), Key("pos-a2a", "a", "a", "set", kw, wp / 2)]
    return keys


def nce_loss(s_pos, s_neg, Z):
    """nce.py:38-58 with a given partition constant Z (python float or 0-d tensor)."""
    K = s_neg.shape[1]
    e_pos, e_neg = torch.exp(s_pos), torch.exp(s_neg)
    c = K * Z
    ln_pmt = -torch.log(e_pos / (e_pos + c)).mean(-1)
    ln_pon = -torch.log(c / (e_neg + c)).sum(-1)
    return (ln_pmt + ln_pon).mean()


def partition_mean(s_neg):
    """nce.py:26 for one rank: mean over (b,k) of exp(s_neg)."""
    return torch.exp(s_neg).mean()


def remap_negatives_avid(raw, y):
    """avid.py:85: raw ~ U[0,N-1) -> skip self."""
    return raw + (raw >= y.view(-1, 1)).long()


def remap_negatives_cma(raw, pos_rows):
    """avid_cma.py:204-207: raw ~ U[0,N-pos_k) -> skip the (sorted) positives of each instance."""
    ref = pos_rows.long() - torch.arange(pos_rows.shape[1], dtype=torch.long).view(1, -1)
    return raw + (raw.unsqueeze(2) >= ref.unsqueeze(1)).sum(2)


def scores(emb_v, emb_a, y, bank_v, bank_a, neg_idx, keys, positive_set=None, T=TEMPERATURE):
    """{key.name: (s_pos (B,P), s_neg (B,K_key))}: avid.py:47-75 / avid_cma.py:150-188.

    Banks are read without grad (they are buffers gathered under no_grad in the reference)."""
    ctx = {"v": l2_normalize(emb_v), "a": l2_normalize(emb_a)}
    mem = {"v": bank_v.detach().to(emb_v.dtype), "a": bank_a.detach().to(emb_v.dtype)}
    out = OrderedDict()
    for k in keys:
        m = mem[k.bank]
        pos_rows = m[y].unsqueeze(1) if k.pos == "self" else m[positive_set[y].long()]
        neg_rows = m[neg_idx[:, :k.num_neg]]
        e = ctx[k.ctx].unsqueeze(2)
        out[k.name] = (torch.bmm(pos_rows, e).squeeze(-1) / T, torch.bmm(neg_rows, e).squeeze(-1) / T)
    return out


def criterion_forward(emb_v, emb_a, y, bank_v, bank_a, neg_idx, keys, Z=-1.0, positive_set=None,
                      rank_partition_means=None):
    """AVID.forward / AVID_CMA.forward up to the total loss.

    Z <= 0 reproduces the first-batch behaviour of nce.py:21-36: Z is the mean of exp(s_neg) of
    the FIRST key scored (mean of `rank_partition_means` + this rank's when given, mimicking the
    all-gather), then frozen and shared by every later key.
    Returns (total, {name: loss}, Z)."""
    sc = scores(emb_v, emb_a, y, bank_v, bank_a, neg_idx, keys, positive_set)
    losses = OrderedDict()
    total = 0.0
    Z = float(Z)
    for k in keys:
        s_pos, s_neg = sc[k.name]
        if Z <= 0:
            with torch.no_grad():
                pm = partition_mean(s_neg)
                if rank_partition_means is not None:
                    pm = torch.stack([pm.to(torch.float32)] + [torch.as_tensor(r, dtype=torch.float32) for r in rank_partition_means]).mean()
                # the reference keeps Z as an fp32 buffer
                Z = float(pm.to(torch.float32))
        losses[k.name] = nce_loss(s_pos, s_neg, Z)
        total = total + k.weight * losses[k.name]
    return total, losses, Z


def criterion_forward_backward(emb_v, emb_a, y, bank_v, bank_a, neg_idx, keys, Z=-1.0, positive_set=None,
                               dtype=torch.float32):
    """Loss and d total / d (un-normalised) embeddings through autograd in `dtype`."""
    ev = emb_v.detach().to(dtype).requires_grad_(True)
    ea = emb_a.detach().to(dtype).requires_grad_(True)
    total, losses, Z = criterion_forward(ev, ea, y, bank_v, bank_a, neg_idx, keys, Z, positive_set)
    total.backward()
    gv = ev.grad if ev.grad is not None else torch.zeros_like(ev)
    ga = ea.grad if ea.grad is not None else torch.zeros_like(ea)
    return {"total": total.detach(), "losses": {k: v.detach() for k, v in losses.items()}, "Z": Z,
            "grad_v": gv, "grad_a": ga}


def bank_update(bank_v, bank_a, emb_v, emb_a, y, momentum=0.5):
    """update_memory (avid.py:103-129) on already-gathered embeddings; in place, returns the banks.

    Embeddings are normalised here (the reference passes the normalised ones, avid.py:78)."""
    mom = momentum if isinstance(momentum, (list, tuple)) else [momentum] * 2
    for bank, emb, m in ((bank_v, emb_v, float(mom[0])), (bank_a, emb_a, float(mom[1]))):
        e = l2_normalize(emb.detach().to(bank.dtype))
        rows = bank[y] * m + e * (1 - m)
        bank[y] = l2_normalize(rows)
    return bank_v, bank_a


def cma_topk(bank_v, bank_a, pos_k, mode="consensus", queries=None, chunk=256):
    """CMASampler.sample_instance (avid_cma.py:42-73) for `queries` (default: all rows).

    similarity over ALL rows, top-(pos_k+1) sorted descending, rank 0 dropped (assumed to be
    the query itself), remaining indices sorted ascending.  Returns (Q, pos_k) int32."""
    N = bank_v.shape[0]
    q = torch.arange(N) if queries is None else torch.as_tensor(queries, dtype=torch.long)
    out = np.zeros((len(q), pos_k), dtype=np.int32)
    for s in range(0, len(q), chunk):
        qi = q[s:s + chunk]
        vs = bank_v @ bank_v[qi].t()
        as_ = bank_a @ bank_a[qi].t()
        if mode == "consensus":
            sim = torch.minimum(vs, as_)
        elif mode == "union":
            sim = torch.maximum(vs, as_)
        elif mode == "video":
            sim = vs
        elif mode == "audio":
            sim = as_
        else:
            raise ValueError(mode)
        idx = torch.topk(sim, pos_k + 1, dim=0, sorted=True)[1][1:].t().numpy()
        out[s:s + len(qi)] = np.sort(idx, axis=1)
    return torch.from_numpy(out)
