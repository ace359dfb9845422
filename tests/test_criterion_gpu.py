"""GPU parity of the criterion path (fused NCE, bank update, samplers, CMA mining) through the C ABI,
against the golden vectors produced by the imported reference and against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import criterion as oc
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, rtol, atol=0.0):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def _make_avid(N, K, xw, momentum, seed):
    from avid_cma_b200.criterions import AVID
    crit = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=momentum, xModal_coeff=xw[0], wModal_coeff=xw[1], device=0)
    crit.nce_average.view1_mem.copy_(synth.bank(N, seed=seed, tag="bank_v"))
    crit.nce_average.view2_mem.copy_(synth.bank(N, seed=seed, tag="bank_a"))
    return crit


@pytest.mark.parametrize("tag", ["cross", "joint", "cfg1"])
def test_avid_matches_reference_golden(golden, tag):
    g = golden("criterion_" + tag)
    N, B, K, seed = int(g["N"]), int(g["B"]), int(g["K"]), int(g["seed"])
    crit = _make_avid(N, K, (float(g["xModal"]), float(g["wModal"])), g["momentum"].tolist(), seed)
    for s in range(int(g["steps"])):
        ev, ea = synth.embeddings(B, seed=seed + 100 * s)
        y = synth.instance_ids(B, N, seed=seed + 100 * s)
        idx = synth.negatives(y, K, N, seed=seed + 100 * s).to(DEV)
        crit.nce_average.sample_negatives = lambda y_, K_, idx=idx: idx
        ev, ea = ev.to(DEV).requires_grad_(True), ea.to(DEV).requires_grad_(True)
        loss, log = crit(ev, ea, y.to(DEV))
        loss.backward()
        _close(crit.criterion.avg_exp_score, g[f"s{s}_Z"], 1e-5)
        _close(loss, g[f"s{s}_total"], 1e-5)
        for k in log:
            _close(log[k], g[f"s{s}_{k}"], 1e-5, 1e-7)
        _close(ev.grad, g[f"s{s}_grad_v"], 1e-3, 1e-7)
        _close(ea.grad, g[f"s{s}_grad_a"], 1e-3, 1e-7)
        _close(crit.nce_average.view1_mem[y.to(DEV)], g[f"s{s}_rows_v"], 1e-5, 1e-7)
        _close(crit.nce_average.view2_mem[y.to(DEV)], g[f"s{s}_rows_a"], 1e-5, 1e-7)


@pytest.mark.parametrize("tag,mode", [("consensus", "consensus"), ("union", "union")])
def test_avid_cma_matches_reference_golden(golden, tag, mode):
    from avid_cma_b200.criterions import AVID_CMA
    g = golden("cma_" + tag)
    N, B, K, pos_k, seed = int(g["N"]), int(g["B"]), int(g["K"]), int(g["pos_k"]), int(g["seed"])
    Kw = None if int(g["Kw"]) < 0 else int(g["Kw"])
    crit = AVID_CMA(num_data=N, embedding_dim=128, num_negatives=K, num_negatives_within=Kw, momentum=0.5,
                    sampling_args={"type": mode, "pos_k": pos_k}, device=0)
    crit.nce_average.view1_mem.copy_(synth.bank(N, seed=seed, tag="bank_v"))
    crit.nce_average.view2_mem.copy_(synth.bank(N, seed=seed, tag="bank_a"))
    crit.nce_average.find_correspondences()
    assert np.array_equal(crit.nce_average.positive_set.cpu().numpy(), g["positive_set"])
    ev, ea = synth.embeddings(B, seed=seed)
    y = synth.instance_ids(B, N, seed=seed)
    neg = torch.from_numpy(g["neg_idx"]).to(DEV)
    crit.nce_average.memory_sampling = lambda y_: (None, neg)
    ev, ea = ev.to(DEV).requires_grad_(True), ea.to(DEV).requires_grad_(True)
    loss, log = crit(ev, ea, y.to(DEV))
    loss.backward()
    _close(crit.criterion.avg_exp_score, g["Z"], 1e-5)
    _close(loss, g["total"], 1e-5)
    for k in log:
        _close(log[k], g[k], 1e-5)
    _close(ev.grad, g["grad_v"], 1e-3, 1e-7)
    _close(ea.grad, g["grad_a"], 1e-3, 1e-7)
    _close(crit.nce_average.view1_mem[y.to(DEV)], g["rows_v"], 1e-5, 1e-7)


def test_device_sampler_support_and_uniformity():
    from avid_cma_b200 import ops
    N, B, K = 1000, 16, 4096
    y = synth.instance_ids(B, N, seed=7).to(DEV)
    idx = ops.sample_negatives(y, K, N, seed=123, offset=0)
    assert idx.min() >= 0 and idx.max() < N
    assert not (idx == y.view(-1, 1)).any()                         # avid.py:85 never returns the instance itself
    counts = torch.bincount(idx.flatten(), minlength=N).double().cpu()
    expected = B * K / (N - 1) * (1 - 0)                              # every other row equally likely
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    assert chi2 < N + 6 * (2 * N) ** 0.5, chi2
    idx2 = ops.sample_negatives(y, K, N, seed=123, offset=B * K)
    assert (idx != idx2).float().mean() > 0.99                        # the offset advances the stream
    assert torch.equal(idx, ops.sample_negatives(y, K, N, seed=123, offset=0))
    # CMA: negatives avoid the (sorted) positive set but may hit the instance itself (avid_cma.py:200-207)
    pos = torch.stack([torch.sort(torch.randperm(N)[:8])[0] for _ in range(N)]).int().to(DEV)
    neg = ops.sample_negatives(y, K, N, seed=5, offset=0, positive_set=pos)
    assert neg.min() >= 0 and neg.max() < N
    hit = (neg.unsqueeze(2) == pos[y].long().unsqueeze(1)).any()
    assert not bool(hit)
    cover = torch.bincount(neg[0], minlength=N).cpu()
    assert int((cover > 0).sum()) > 0.95 * (N - 8)


def test_device_sampler_matches_host_philox():
    """The kernel's draw is the documented Philox4x32-10 stream: recompute it in numpy."""
    from avid_cma_b200 import ops

    def philox(seed, ctr):
        M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
        c = [ctr & 0xFFFFFFFF, ctr >> 32, 0, 0]
        k0, k1 = seed & 0xFFFFFFFF, seed >> 32
        for _ in range(10):
            p0, p1 = M0 * c[0], M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k1, p0 & 0xFFFFFFFF]
            k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
        return c

    N, B, K, seed, off = 5000, 3, 7, 0x1234567890ABCDEF, 99
    y = torch.tensor([0, 2500, 4999], device=DEV)
    got = ops.sample_negatives(y, K, N, seed=seed, offset=off).cpu()
    for b in range(B):
        for k in range(K):
            r = philox(seed, off + b * K + k)
            u = (((r[1] << 32) | r[0]) * (N - 1)) >> 64
            assert int(got[b, k]) == u + (1 if u >= int(y[b]) else 0)


def test_fused_in_kernel_sampling_equals_standalone_sampler():
    """Without injected indices the NCE kernel draws its own negatives; it must use exactly the indices
    avid_sample_negatives returns for the same (seed, offset), and report them through neg_idx_out."""
    from avid_cma_b200 import ops
    N, B, K = 4096, 8, 256
    bv, ba = synth.bank(N, seed=11, tag="bank_v").to(DEV), synth.bank(N, seed=11, tag="bank_a").to(DEV)
    ev, ea = [t.to(DEV) for t in synth.embeddings(B, seed=11)]
    y = synth.instance_ids(B, N, seed=11).to(DEV)
    keys = [(0, 1, 0, K, 0.5), (1, 0, 0, K, 0.5)]
    Z = torch.tensor(2.0, device=DEV)
    ws = ops.nce_workspace(B, K, 0, 2, DEV)
    res = []
    for inject in (False, True):
        out = [torch.empty(2, device=DEV), torch.empty(1, device=DEV), torch.empty(B, 128, device=DEV), torch.empty(B, 128, device=DEV)]
        used = torch.empty(B, K, dtype=torch.int64, device=DEV)
        neg = ops.sample_negatives(y, K, N, seed=77, offset=1000) if inject else None
        a = ops.make_nce_args(ev, ea, y, bv, ba, keys, K, Z, neg_idx=neg, seed=77, offset=1000, loss_keys=out[0], loss_total=out[1],
                              grad_v=out[2], grad_a=out[3], neg_idx_out=used)
        ops.nce_forward_backward(a, ws)
        res.append((out, used))
    assert torch.equal(res[0][1], res[1][1])
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    r = oc.criterion_forward_backward(ev.cpu(), ea.cpu(), y.cpu(), bv.cpu(), ba.cpu(), res[0][1].cpu(), oc.avid_keys(K), 2.0, dtype=torch.float64)
    _close(res[0][0][1], r["total"], 1e-5)
    _close(res[0][0][2], r["grad_v"], 1e-4, 1e-7)


@pytest.mark.parametrize("N,B,K", [(240000, 64, 1024), (240000, 64, 4096), (50000, 33, 100)])
def test_fused_nce_full_size_vs_fp64_oracle(N, B, K):
    """BASELINE config 2 / 5 sizes (ragged third case): loss, per-key losses, gradients and scores vs the fp64 oracle."""
    from avid_cma_b200 import ops
    bv, ba = synth.bank(N, seed=21, tag="bank_v"), synth.bank(N, seed=21, tag="bank_a")
    ev, ea = synth.embeddings(B, seed=21)
    y = synth.instance_ids(B, N, seed=21)
    idx = synth.negatives(y, K, N, seed=21)
    keys = oc.avid_keys(K, 1.0, 1.0)
    kt = [({"v": 0, "a": 1}[k.ctx], {"v": 0, "a": 1}[k.bank], 0, k.num_neg, k.weight) for k in keys]
    d = lambda t: t.to(DEV)
    Z = torch.tensor(1.7, device=DEV)
    out = [torch.empty(4, device=DEV), torch.empty(1, device=DEV), torch.empty(B, 128, device=DEV), torch.empty(B, 128, device=DEV)]
    scores = torch.full((4, B, 1 + K), float("nan"), device=DEV)
    a = ops.make_nce_args(d(ev), d(ea), d(y), d(bv), d(ba), kt, K, Z, neg_idx=d(idx), loss_keys=out[0], loss_total=out[1],
                          grad_v=out[2], grad_a=out[3], scores=scores)
    ops.nce_forward_backward(a, ops.nce_workspace(B, K, 0, 4, DEV))
    r = oc.criterion_forward_backward(ev, ea, y, bv, ba, idx, keys, 1.7, dtype=torch.float64)
    _close(out[1], r["total"], 2e-6)
    for i, k in enumerate(keys):
        _close(out[0][i], r["losses"][k.name], 2e-6)
    _close(out[2], r["grad_v"], 1e-4, 1e-8)
    _close(out[3], r["grad_a"], 1e-4, 1e-8)
    sc = oc.scores(ev.double(), ea.double(), y, bv, ba, idx, keys)
    for i, k in enumerate(keys):
        _close(scores[i, :, :1], sc[k.name][0], 1e-5, 1e-5)
        _close(scores[i, :, 1:], sc[k.name][1], 1e-5, 1e-5)


def test_partition_function_first_batch():
    from avid_cma_b200 import ops
    N, B, K = 3000, 9, 200
    bv, ba = synth.bank(N, seed=31, tag="bank_v"), synth.bank(N, seed=31, tag="bank_a")
    ev, ea = synth.embeddings(B, seed=31)
    y = synth.instance_ids(B, N, seed=31)
    idx = synth.negatives(y, K, N, seed=31)
    d = lambda t: t.to(DEV)
    a = ops.make_nce_args(d(ev), d(ea), d(y), d(bv), d(ba), [(0, 1, 0, K, 0.5), (1, 0, 0, K - 50, 0.5)], K, None, neg_idx=d(idx))
    ws = ops.nce_workspace(B, K, 0, 2, DEV)
    z = torch.empty(1, device=DEV)
    sc = oc.scores(ev.double(), ea.double(), y, bv, ba, idx, [oc.Key("v2a", "v", "a", "self", K, .5), oc.Key("a2v", "a", "v", "self", K - 50, .5)])
    ops.nce_partition_mean(a, 0, z, ws)
    _close(z, oc.partition_mean(sc["v2a"][1]), 1e-5)
    ops.nce_partition_mean(a, 1, z, ws)
    _close(z, oc.partition_mean(sc["a2v"][1]), 1e-5)


def test_sharded_scoring_sums_to_unsharded():
    """every rank scores all queries against the rows it owns; the summed partials, finalised,
    equal the replicated-bank result.  Two 'ranks' emulated on one GPU."""
    from avid_cma_b200 import ops
    N, B, K = 10001, 12, 300
    bv, ba = synth.bank(N, seed=41, tag="bank_v").to(DEV), synth.bank(N, seed=41, tag="bank_a").to(DEV)
    ev, ea = [t.to(DEV) for t in synth.embeddings(B, seed=41)]
    y = synth.instance_ids(B, N, seed=41).to(DEV)
    keys = [(0, 1, 0, K, 0.5), (1, 0, 0, K, 0.5)]
    Z = torch.tensor(2.2, device=DEV)
    ws = ops.nce_workspace(B, K, 0, 2, DEV)
    ref = [torch.empty(2, device=DEV), torch.empty(1, device=DEV), torch.empty(B, 128, device=DEV), torch.empty(B, 128, device=DEV)]
    a = ops.make_nce_args(ev, ea, y, bv, ba, keys, K, Z, seed=9, offset=5, loss_keys=ref[0], loss_total=ref[1], grad_v=ref[2], grad_a=ref[3])
    ops.nce_forward_backward(a, ws)
    gh_v, gh_a, lp = torch.zeros(B, 128, device=DEV), torch.zeros(B, 128, device=DEV), torch.zeros(2, B, device=DEV)
    cut = 4321
    z_parts = []
    for lo, hi in ((0, cut), (cut, N)):
        pv, pa, pl = torch.empty(B, 128, device=DEV), torch.empty(B, 128, device=DEV), torch.empty(2, B, device=DEV)
        a = ops.make_nce_args(ev, ea, y, bv[lo:hi].contiguous(), ba[lo:hi].contiguous(), keys, K, Z, num_rows=N, row_begin=lo, row_end=hi,
                              seed=9, offset=5, grad_hat_v=pv, grad_hat_a=pa, loss_part=pl)
        ops.nce_forward_backward(a, ws)
        gh_v += pv; gh_a += pa; lp += pl
        zp = torch.empty(1, device=DEV)
        ops.nce_partition_mean(a, 0, zp, ws)
        z_parts.append(zp)
    out = [torch.empty(2, device=DEV), torch.empty(1, device=DEV), torch.empty(B, 128, device=DEV), torch.empty(B, 128, device=DEV)]
    a = ops.make_nce_args(ev, ea, y, bv, ba, keys, K, Z, loss_keys=out[0], loss_total=out[1], grad_v=out[2], grad_a=out[3],
                          grad_hat_v=gh_v, grad_hat_a=gh_a, loss_part=lp)
    ops.nce_finalize(a, ws)
    for o, r in zip(out, ref):
        _close(o, r, 1e-5, 1e-8)
    # partition function: partial sums over the shards / (B*K) == unsharded mean
    a = ops.make_nce_args(ev, ea, y, bv, ba, keys, K, None, seed=9, offset=5)
    zf = torch.empty(1, device=DEV)
    ops.nce_partition_mean(a, 0, zf, ws)
    _close((z_parts[0] + z_parts[1]) / (B * K), zf, 1e-5)


def test_bank_update_and_normalize_kernels():
    from avid_cma_b200 import ops
    N, B = 777, 50
    bv, ba = synth.bank(N, seed=51, tag="bank_v"), synth.bank(N, seed=51, tag="bank_a")
    ev, ea = synth.embeddings(B, seed=51)
    y = synth.instance_ids(B, N, seed=51)
    gv, ga = bv.to(DEV), ba.to(DEV)
    ops.bank_update(gv, ga, ev.to(DEV), ea.to(DEV), y.to(DEV), 0.3, 0.9)
    oc.bank_update(bv, ba, ev, ea, y, [0.3, 0.9])
    _close(gv, bv, 1e-5, 1e-7)
    _close(ga, ba, 1e-5, 1e-7)
    # sharded update touches only owned rows
    g2 = bv.to(DEV)[100:400].contiguous()
    before = g2.clone()
    ops.bank_update(g2, ba.to(DEV)[100:400].contiguous(), ev.to(DEV), ea.to(DEV), y.to(DEV), 0.5, 0.5, row_begin=100, row_end=400)
    owned = ((y >= 100) & (y < 400))
    changed = (g2 != before).any(1).cpu()
    assert int(changed.sum()) == int(owned.sum())
    x = torch.randn(1000, 128, device=DEV) * 3
    x[5] = 0                                                         # zero row: x / max(0, eps) = 0
    ref = oc.l2_normalize(x.cpu().double())
    _close(ops.rows_l2_normalize_(x), ref, 1e-6, 1e-7)


def _sim64(bv, ba, q, mode):
    """fp64 similarity of query row q against every row (avid_cma.py:52-64)."""
    sv, sa = bv.double() @ bv[q].double(), ba.double() @ ba[q].double()
    return {"consensus": torch.minimum(sv, sa), "union": torch.maximum(sv, sa), "video": sv, "audio": sa}[mode].cpu().numpy()


def _assert_equal_up_to_proven_ties(got, want, bv, ba, mode, pos_k, query_rows=None, tol=1e-6):
    """Index outputs must be identical.  The only admissible difference between two correct top-k searches that sum in a
    different order (fp32 kernel vs fp64 oracle, tensor-core candidates + fp32 re-score vs fp32 scan) is WHICH of several
    candidates tied at the k-th boundary is kept (or, with near-duplicate rows, which of two candidates tied for rank 0 is the one
    the reference drops as "the query itself", avid_cma.py:69): every differing row is checked to be exactly that -- all indices in
    the symmetric difference have an exact (fp64) similarity within `tol` of the boundary value or of the best value.  Returns the
    number of such rows."""
    got, want = np.asarray(got), np.asarray(want)
    rows = np.nonzero((got != want).any(1))[0]
    for r in rows:
        q = int(r if query_rows is None else query_rows[r])
        sim = _sim64(bv, ba, q, mode)
        boundary, best = np.sort(sim)[-(pos_k + 1)], sim.max()       # the (pos_k + 1)-th best includes the query itself
        diff = set(got[r].tolist()) ^ set(want[r].tolist())
        assert diff and all(0 <= j < sim.shape[0] and min(abs(sim[j] - boundary), abs(sim[j] - best)) < tol for j in diff), (r, sorted(diff), boundary,
                                                                                                 [float(sim[j]) for j in diff if 0 <= j < sim.shape[0]])
    assert len(rows) <= max(1, got.shape[0] // 50), len(rows)        # and ties are rare for random unit rows
    return len(rows)


@pytest.mark.parametrize("mode", ["consensus", "union", "video", "audio"])
def test_cma_topk_vs_oracle(mode):
    """Ragged sizes (queries and candidates not multiples of the 64-row tiles), candidates fed in two shards."""
    from avid_cma_b200 import ops
    N, pos_k = 1000 + 37, 32
    bv, ba = synth.bank(N, seed=61, tag="bank_v"), synth.bank(N, seed=61, tag="bank_a")
    want = oc.cma_topk(bv.double(), ba.double(), pos_k, mode)
    gv, ga = bv.to(DEV), ba.to(DEV)
    cut = 300
    got = ops.cma_topk(gv, ga, [(gv[:cut].contiguous(), ga[:cut].contiguous(), 0), (gv[cut:].contiguous(), ga[cut:].contiguous(), cut)], pos_k, mode)
    got, want = got.cpu().numpy(), want.numpy()
    _assert_equal_up_to_proven_ties(got, want, bv, ba, mode, pos_k)
    sub = ops.cma_topk(gv[500:563].contiguous(), ga[500:563].contiguous(), [(gv, ga, 0)], pos_k, mode).cpu().numpy()
    assert (sub == got[500:563]).all()


@pytest.mark.parametrize("mode", ["consensus", "union", "video", "audio"])
def test_cma_tensor_core_path_equals_fp32_path(mode):
    """Tensor-core candidate generation + exact re-scoring + certificate (csrc/cma_tc.cu) returns the positive sets of the
    fp32 kernel (csrc/cma.cu); every query is certified at the rigorous eps, and when the certificate is made to fail
    (huge eps) the re-mining fallback returns the same sets."""
    from avid_cma_b200 import ops
    N, pos_k = 2000 + 77, 32
    bv, ba = synth.bank(N, seed=62, tag="bank_v").to(DEV), synth.bank(N, seed=62, tag="bank_a").to(DEV)
    cut = 1000 + 13
    shards = [(bv[:cut].contiguous(), ba[:cut].contiguous(), 0), (bv[cut:].contiguous(), ba[cut:].contiguous(), cut)]
    exact = ops.cma_topk(bv, ba, shards, pos_k, mode, exact=True)
    st = {}
    tc = ops.cma_topk(bv, ba, shards, pos_k, mode, exact=False, stats=st)
    assert st["uncertified"] == 0
    _assert_equal_up_to_proven_ties(tc.cpu().numpy(), exact.cpu().numpy(), bv.cpu(), ba.cpu(), mode, pos_k)
    st2 = {}
    fb = ops.cma_topk(bv, ba, shards, pos_k, mode, exact=False, eps=10.0, stats=st2)
    assert st2["uncertified"] == N
    assert torch.equal(fb, exact)
    # planted near-duplicates: rows 5 and 1500 almost equal -> each is the other's best positive
    bv2, ba2 = bv.clone(), ba.clone()
    bv2[1500] = torch.nn.functional.normalize(bv2[5] + 1e-3 * bv2[1500], dim=0)
    ba2[1500] = torch.nn.functional.normalize(ba2[5] + 1e-3 * ba2[1500], dim=0)
    tc2 = ops.cma_topk(bv2, ba2, [(bv2, ba2, 0)], pos_k, mode, exact=False)
    ex2 = ops.cma_topk(bv2, ba2, [(bv2, ba2, 0)], pos_k, mode, exact=True)
    _assert_equal_up_to_proven_ties(tc2.cpu().numpy(), ex2.cpu().numpy(), bv2.cpu(), ba2.cpu(), mode, pos_k)
    assert 1500 in tc2[5].tolist() or 5 in tc2[5].tolist()



@pytest.mark.parametrize("mode", ["video", "audio"])
def test_cma_single_modality_mining_matches_reference_golden(golden, mode):
    """Tensor-core mining with one modality (only that modality's tiles are loaded and multiplied) == the imported reference."""
    from avid_cma_b200 import ops
    g = golden("cma_mining_" + mode)
    N, pos_k, seed = int(g["N"]), int(g["pos_k"]), int(g["seed"])
    bv, ba = synth.bank(N, seed=seed, tag="bank_v").to(DEV), synth.bank(N, seed=seed, tag="bank_a").to(DEV)
    got = ops.cma_topk(bv, ba, [(bv, ba, 0)], pos_k, mode).cpu().numpy()
    _assert_equal_up_to_proven_ties(got, g["positive_set"], bv.cpu(), ba.cpu(), mode, pos_k)


def test_cma_tensor_core_equals_fp32_at_240k_rows():
    """BASELINE config 4 size: on a 240 k-row bank the tensor-core path (fp16 candidates on tcgen05 -> exact fp32 re-score ->
    certificate) returns the positive sets of the fp32 CUDA-core scan for a 2 k-query slice; all queries certified."""
    from avid_cma_b200 import ops
    N, pos_k, nq = 240000, 32, 2048
    g = torch.Generator(device=DEV).manual_seed(240)
    bv = ops.rows_l2_normalize_(torch.randn(N, 128, device=DEV, generator=g))
    ba = ops.rows_l2_normalize_(torch.randn(N, 128, device=DEV, generator=g))
    # planted clusters so that the top-32 is not only noise: 64 groups of 40 rows around a common centre
    for c in range(64):
        rows = torch.arange(c * 3000, c * 3000 + 40, device=DEV)
        bv[rows] = torch.nn.functional.normalize(bv[rows[0]] + 0.6 * bv[rows], dim=1)
        ba[rows] = torch.nn.functional.normalize(ba[rows[0]] + 0.6 * ba[rows], dim=1)
    q_rows = torch.cat([torch.arange(0, 1024, device=DEV), torch.arange(3000 * 7, 3000 * 7 + 512, device=DEV),
                        torch.arange(N - 512, N, device=DEV)])
    qv, qa = bv[q_rows].contiguous(), ba[q_rows].contiguous()
    st = {}
    tc = ops.cma_topk(qv, qa, [(bv, ba, 0)], pos_k, "consensus", exact=False, stats=st)
    ex = ops.cma_topk(qv, qa, [(bv, ba, 0)], pos_k, "consensus", exact=True)
    assert st["uncertified"] == 0
    assert tc.shape == (nq, pos_k)
    _assert_equal_up_to_proven_ties(tc.cpu().numpy(), ex.cpu().numpy(), bv.cpu(), ba.cpu(), "consensus", pos_k, query_rows=q_rows.cpu().numpy())
    # a planted cluster really is found: the positives of a cluster row are its cluster mates
    members = set(range(3000 * 7, 3000 * 7 + 40))
    row = tc[1024 + 3].tolist()
    assert len(members & set(row)) >= 30, row
