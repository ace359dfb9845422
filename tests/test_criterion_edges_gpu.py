"""Edge cases of the criterion path the reference can run into: repeated instance ids inside one batch
(clips_per_video > 1: video_db.py:98), minimal sizes, negative counts that are not multiples of the kernel's chunking, a zero
embedding (F.normalize's eps clamp), and the error behaviour of the drop-in classes."""
import pytest
import torch

from oracle import criterion as oc
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(ev, ea, y, bv, ba, idx, keys, Z):
    from avid_cma_b200 import ops
    B, K = idx.shape
    kt = [({"v": 0, "a": 1}[k.ctx], {"v": 0, "a": 1}[k.bank], 0, k.num_neg, k.weight) for k in keys]
    d = lambda t: t.to(DEV)
    out = [torch.empty(len(keys), device=DEV), torch.empty(1, device=DEV), torch.empty(B, 128, device=DEV), torch.empty(B, 128, device=DEV)]
    a = ops.make_nce_args(d(ev), d(ea), d(y), d(bv), d(ba), kt, K, torch.tensor(Z, device=DEV), neg_idx=d(idx), loss_keys=out[0],
                          loss_total=out[1], grad_v=out[2], grad_a=out[3])
    ws = ops.nce_workspace(B, K, 0, len(keys), DEV)
    ops.nce_forward_backward(a, ws)
    first = [o.clone() for o in out]
    ops.nce_forward_backward(a, ws)                      # the workspace's ticket counters were left zero: same result again
    torch.cuda.synchronize()
    for x, z in zip(first, out):
        assert torch.equal(x, z)
    return out


@pytest.mark.parametrize("N,B,K", [(3, 1, 1), (40, 2, 7), (500, 5, 129), (1000, 64, 33)])
def test_small_and_ragged_sizes_vs_fp64_oracle(N, B, K):
    bv, ba = synth.bank(N, seed=5, tag="bank_v"), synth.bank(N, seed=5, tag="bank_a")
    ev, ea = synth.embeddings(B, seed=5)
    y = synth.instance_ids(B, N, seed=5)
    idx = synth.negatives(y, K, N, seed=5)
    keys = oc.avid_keys(K, 1.0, 1.0)
    out = _run(ev, ea, y, bv, ba, idx, keys, 1.3)
    r = oc.criterion_forward_backward(ev, ea, y, bv, ba, idx, keys, 1.3, dtype=torch.float64)
    torch.testing.assert_close(out[1].cpu().double().squeeze(), torch.as_tensor(float(r["total"]), dtype=torch.float64), rtol=5e-6, atol=0)
    torch.testing.assert_close(out[2].cpu().double(), r["grad_v"].double(), rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(out[3].cpu().double(), r["grad_a"].double(), rtol=1e-4, atol=1e-8)


def test_repeated_instance_ids_in_one_batch():
    """Two clips of the same video in a batch (same y): the loss treats them independently; the bank update has one winner per
    row (avid.py:119-129 index_copy_ with duplicates is 'unspecified winner' in the reference too)."""
    from avid_cma_b200 import ops
    N, B, K = 300, 6, 64
    bv, ba = synth.bank(N, seed=9, tag="bank_v"), synth.bank(N, seed=9, tag="bank_a")
    ev, ea = synth.embeddings(B, seed=9)
    y = torch.tensor([7, 120, 7, 33, 120, 7])
    idx = synth.negatives(y, K, N, seed=9)
    keys = oc.avid_keys(K)
    out = _run(ev, ea, y, bv, ba, idx, keys, 2.0)
    r = oc.criterion_forward_backward(ev, ea, y, bv, ba, idx, keys, 2.0, dtype=torch.float64)
    torch.testing.assert_close(out[2].cpu().double(), r["grad_v"].double(), rtol=1e-4, atol=1e-8)
    gv, ga = bv.to(DEV).clone(), ba.to(DEV).clone()
    ops.bank_update(gv, ga, ev.to(DEV), ea.to(DEV), y.to(DEV), 0.5, 0.5)
    torch.cuda.synchronize()
    nv = torch.nn.functional.normalize(ev.double(), dim=1)
    for row in (7, 120, 33):
        cands = [torch.nn.functional.normalize(0.5 * bv[row].double() + 0.5 * nv[b], dim=0) for b in range(B) if int(y[b]) == row]
        got = gv[row].cpu().double()
        # every element comes from one of the batch's candidates for this row (like index_copy_ with duplicates, the winner is
        # unspecified -- and, per 16-byte chunk, so it is here)
        err = torch.stack([(got - c).abs() for c in cands]).min(0).values
        assert float(err.max()) < 1e-6
    untouched = torch.ones(N, dtype=torch.bool)
    untouched[y] = False
    assert torch.equal(gv.cpu()[untouched], bv[untouched])


def test_zero_embedding_row_is_clamped_like_f_normalize():
    N, B, K = 200, 3, 32
    bv, ba = synth.bank(N, seed=11, tag="bank_v"), synth.bank(N, seed=11, tag="bank_a")
    ev, ea = synth.embeddings(B, seed=11)
    ev[1].zero_()                                            # x / max(||x||, 1e-12) = 0: scores 0, finite loss, finite gradient
    y = synth.instance_ids(B, N, seed=11)
    idx = synth.negatives(y, K, N, seed=11)
    keys = oc.avid_keys(K)
    out = _run(ev, ea, y, bv, ba, idx, keys, 2.0)
    r = oc.criterion_forward(ev, ea, y, bv, ba, idx, keys, 2.0)
    assert torch.isfinite(out[1]).all() and torch.isfinite(out[2]).all() and torch.isfinite(out[3]).all()
    torch.testing.assert_close(out[1].cpu().squeeze(), torch.as_tensor(float(r[0])), rtol=1e-5, atol=0)


def test_drop_in_classes_reject_what_they_cannot_run():
    from avid_cma_b200.criterions import AVID
    from avid_cma_b200 import models
    with pytest.raises((ValueError, AssertionError, RuntimeError)):
        AVID(num_data=100, embedding_dim=64, num_negatives=16, device=0)          # the kernels are specialised for D = 128
    crit = AVID(num_data=100, embedding_dim=128, num_negatives=16, momentum=0.5, device=0)
    with pytest.raises(RuntimeError):
        crit(torch.randn(2, 128), torch.randn(2, 128), torch.tensor([1, 2]))       # CPU tensors: there is no CPU path
    tower = models.R2Plus1D(depth=18).to(DEV)
    with pytest.raises(RuntimeError):
        tower(torch.randn(1, 3, 4, 32, 32))
