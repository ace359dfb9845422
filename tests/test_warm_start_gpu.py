"""Warm start from another run's checkpoint: the `checkpoint=` argument of `av_wrapper` (av_wrapper.py:72-74,
`module.`-prefixed keys), of `AVID` (avid.py:187-200) and of `AVID_CMA` (avid_cma.py:308-319: banks and Z restored BEFORE the
positives are mined), with the reference's checkpoint layout -- including a shape-[1] `avg_exp_score`, which is what a
single-process reference run saves (nce.py:35), and several `*avg_exp_score*` entries that are averaged (avid.py:196)."""
import numpy as np
import pytest
import torch

from oracle import criterion as oc
from oracle import synth, towers

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N, K = 300, 64


def _reference_layout_checkpoint(path, z_entries):
    model_sd = synth.fill_state_dict(towers.state_dict_template(), seed=21)
    crit_sd = {"nce_average.view1_mem": synth.bank(N, seed=21, tag="bank_v"), "nce_average.view2_mem": synth.bank(N, seed=21, tag="bank_a")}
    crit_sd.update(z_entries)
    torch.save({"epoch": 7, "model": {"module." + k: v for k, v in model_sd.items()}, "optimizer": {},
                "train_criterion": crit_sd}, path)
    return model_sd, crit_sd


def test_av_wrapper_checkpoint_argument(tmp_path):
    from avid_cma_b200 import models
    path = str(tmp_path / "ckp.pth.tar")
    model_sd, _ = _reference_layout_checkpoint(path, {"criterion.avg_exp_score": torch.tensor([2.5])})
    model = models.av_wrapper('R2Plus1D', {'depth': 18}, 'Conv2D', {'depth': 10}, proj_dim=[512, 512, 128], checkpoint=path)
    got = model.state_dict()
    assert list(got.keys()) == list(model_sd.keys())
    for k in model_sd:
        assert torch.equal(got[k].cpu(), model_sd[k]), k


@pytest.mark.parametrize("z_entries,want_z", [
    ({"criterion.avg_exp_score": torch.tensor([2.5])}, 2.5),                                                 # shape [1]: single-process run
    ({"criterion.avg_exp_score": torch.tensor(3.0)}, 3.0),                                                   # shape []: distributed run
    ({"criterion.avg_exp_score": torch.tensor([2.0]), "criterion2.avg_exp_score": torch.tensor(4.0)}, 3.0),  # all entries are averaged
])
def test_avid_warm_start_restores_banks_and_partition(tmp_path, z_entries, want_z):
    from avid_cma_b200.criterions import AVID
    path = str(tmp_path / "ckp.pth.tar")
    _, crit_sd = _reference_layout_checkpoint(path, z_entries)
    torch.manual_seed(3)
    crit = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5, xModal_coeff=1., wModal_coeff=0., checkpoint=path, device=0)
    assert torch.equal(crit.nce_average.view1_mem.cpu(), crit_sd["nce_average.view1_mem"])
    assert torch.equal(crit.nce_average.view2_mem.cpu(), crit_sd["nce_average.view2_mem"])
    assert crit.criterion.avg_exp_score.shape == () and float(crit.criterion.avg_exp_score) == want_z
    # the restored Z is FROZEN: the first batch does not re-estimate it (nce.py:21-24), and the loss is the oracle's at that Z
    ev, ea = synth.embeddings(8, seed=4)
    y = synth.instance_ids(8, N, seed=4)
    idx = synth.negatives(y, K, N, seed=4)
    crit.nce_average.sample_negatives = lambda y_, K_: idx.to(DEV)
    loss, _ = crit(ev.to(DEV), ea.to(DEV), y.to(DEV))
    ref = oc.criterion_forward_backward(ev, ea, y, crit_sd["nce_average.view1_mem"], crit_sd["nce_average.view2_mem"], idx, oc.avid_keys(K), want_z)
    assert float(crit.criterion.avg_exp_score) == want_z
    np.testing.assert_allclose(float(loss), float(ref["total"]), rtol=1e-5)
    # the state_dict keeps the reference's keys (+ our sampler position) and round-trips
    sd = crit.state_dict()
    assert {"nce_average.view1_mem", "nce_average.view2_mem", "criterion.avg_exp_score"} <= set(sd)
    crit2 = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=0)
    crit2.load_state_dict(sd)
    assert torch.equal(crit2.nce_average.view1_mem, crit.nce_average.view1_mem)
    assert (crit2.nce_average._seed, crit2.nce_average._offset) == (crit.nce_average._seed, crit.nce_average._offset)


def test_avid_cma_warm_start_mines_on_the_restored_banks(tmp_path):
    from avid_cma_b200.criterions import AVID_CMA
    path = str(tmp_path / "ckp.pth.tar")
    _, crit_sd = _reference_layout_checkpoint(path, {"criterion.avg_exp_score": torch.tensor([1.75])})
    crit = AVID_CMA(num_data=N, embedding_dim=128, num_negatives=K, num_negatives_within=16, momentum=0.5,
                    sampling_args={"type": "consensus", "pos_k": 8}, checkpoint=path, device=0)
    assert float(crit.criterion.avg_exp_score) == 1.75
    want = oc.cma_topk(crit_sd["nce_average.view1_mem"].double(), crit_sd["nce_average.view2_mem"].double(), 8, "consensus")
    assert torch.equal(crit.nce_average.positive_set.cpu(), want)      # mined AFTER the restore (avid_cma.py:305-323)
    assert "nce_average.positive_set" in crit.state_dict()


def test_out_of_range_instance_index_raises_like_the_reference():
    """avid.py:57-58 indexes the bank with y: an index >= num_data raises IndexError in the reference.  Here the kernel flags it
    (nothing is read or written out of bounds) and the criterion raises on its next call."""
    from avid_cma_b200.criterions import AVID
    crit = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5, xModal_coeff=1., wModal_coeff=0., device=0)
    before = crit.nce_average.view1_mem.clone()
    ev, ea = synth.embeddings(4, seed=5)
    y = torch.tensor([1, N + 5, 7, -3])
    loss, _ = crit(ev.to(DEV), ea.to(DEV), y.to(DEV))
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    changed = (crit.nce_average.view1_mem != before).any(1).nonzero().flatten().tolist()
    assert changed == [1, 7]
    with pytest.raises(IndexError):
        crit(ev.to(DEV), ea.to(DEV), torch.tensor([1, 2, 3, 4], device=DEV))
    crit(ev.to(DEV), ea.to(DEV), torch.tensor([1, 2, 3, 4], device=DEV))      # the flag is cleared once reported


def test_bank_update_with_duplicate_indices_keeps_one_complete_update():
    """Duplicate instance ids in a gathered batch (clips_per_video > 1, sampler padding): exactly one complete, unit-norm update
    per row -- the LAST occurrence (index_copy_ semantics, avid.py:119-129), never a torn row."""
    from avid_cma_b200 import ops
    bv, ba = synth.bank(N, seed=8, tag="bank_v"), synth.bank(N, seed=8, tag="bank_a")
    ev, ea = synth.embeddings(96, seed=8)
    y = torch.from_numpy(np.random.RandomState(8).randint(0, 12, size=96).astype(np.int64))       # 96 updates on 12 rows
    gv, ga = bv.to(DEV), ba.to(DEV)
    ops.bank_update(gv, ga, ev.to(DEV), ea.to(DEV), y.to(DEV), 0.5, 0.5)
    last = {int(r): i for i, r in enumerate(y.tolist())}
    keep = torch.tensor(sorted(last.values()))
    wv, wa = bv.clone(), ba.clone()
    oc.bank_update(wv, wa, ev[keep], ea[keep], y[keep], 0.5)
    np.testing.assert_allclose(gv.cpu().numpy(), wv.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ga.cpu().numpy(), wa.numpy(), rtol=1e-5, atol=1e-6)


def test_bank_init_rows_are_a_function_of_seed_and_row():
    """avid_bank_init: unit-norm N(0,1)-direction rows; a shard filled on its own equals the same rows of the full bank."""
    from avid_cma_b200 import ops
    full = ops.bank_init_(torch.empty(5000, 128, device=DEV), 0, 1234, 0)
    part = ops.bank_init_(torch.empty(700, 128, device=DEV), 2100, 1234, 0)
    other = ops.bank_init_(torch.empty(5000, 128, device=DEV), 0, 1234, 1)
    assert torch.equal(part, full[2100:2800])
    assert not torch.equal(other, full)
    np.testing.assert_allclose(full.norm(dim=1).cpu().numpy(), 1.0, rtol=1e-5)
    x = full.double() * (128 ** 0.5)                 # components of a uniformly random direction: mean 0, variance ~1
    assert abs(float(x.mean())) < 0.01 and abs(float(x.var()) - 1.0) < 0.02
    cos = (full[:2000] @ full[2000:4000].t()).flatten()
    assert abs(float(cos.mean())) < 1e-3 and abs(float(cos.std()) - 128 ** -0.5) < 2e-3
