"""The N > 1 criterion path: row-partitioned memory banks on two ranks.

CPU (`gloo`, world_size 2, runs everywhere): the HOST protocol of criterions/avid.py::_run_sharded -- gather order, Philox /
injected-negative bookkeeping, partial reduction, per-rank finalize, owner-only bank update, first-step partition function --
with the CUDA entry points replaced by torch-CPU stand-ins that restate what each kernel computes for ONE shard (test
infrastructure, like oracle/).  The result must equal the unsharded oracle on the full banks.

GPU (`nccl`, needs 2 devices, `-m gpu`): the same check through the real kernels.
"""
import os
import socket
import sys
from types import SimpleNamespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, B, K, T = 96, 4, 40, 0.07          # 96 rows over 2 ranks; ranks' y and negatives straddle both shards


# ---------------------------------------------------------------------------------------------- torch-CPU kernel stand-ins
def _normalize(x):
    return x / x.pow(2).sum(1, keepdim=True).sqrt().clamp_min(1e-12)


def _dense(t, tail):
    """Packed per-rank views (W, B, ...) of the step's gathered buffer -> dense (W*B, ...) copies; dense tensors pass through."""
    return t.reshape((-1,) + tuple(tail)) if t is not None else None


def _fake_make_nce_args(emb_v, emb_a, y, bank_v, bank_a, keys, num_neg, Z, *, num_rows=None, row_begin=0, row_end=None, neg_idx=None,
                        seed=0, offset=0, positive_set=None, mean_batch=0, temperature=0.07, bad_index=None, **out):
    num_rows = bank_v.shape[0] if num_rows is None else num_rows
    grouped = emb_v.dim() == 3
    if grouped:       # the kernel addresses the rank records by stride; the stand-in copies them out
        assert emb_v.stride(0) == emb_a.stride(0) == 2 * y.stride(0) and emb_v[0].is_contiguous() and y[0].is_contiguous()
    emb_v, emb_a, y, neg_idx = _dense(emb_v, (128,)), _dense(emb_a, (128,)), _dense(y, ()), _dense(neg_idx, (num_neg,))
    return SimpleNamespace(emb=(emb_v, emb_a), y=y, bank=(bank_v, bank_a), keys=keys, K=num_neg, Z=Z, N=num_rows, row_begin=row_begin,
                           row_end=num_rows if row_end is None else row_end, neg_idx=neg_idx, mean_batch=mean_batch or emb_v.shape[0],
                           T=temperature, out=out, grouped=grouped)


def _partial_terms(a, with_loss=True):
    """Scores of the rows this shard holds, for every query of the call; returns (per-key masked terms, ehat leaves)."""
    assert a.neg_idx is not None, "the CPU stand-in needs injected negatives"
    ehat = [_normalize(e.detach()).requires_grad_(True) for e in a.emb]
    terms = []
    for ctx, bank, pos_mode, kn, w in a.keys:
        assert pos_mode == 0
        idx = torch.cat([a.y.view(-1, 1), a.neg_idx[:, :kn]], 1)                    # (Bq, 1 + kn)
        held = (idx >= a.row_begin) & (idx < a.row_end)
        rows = a.bank[bank][(idx - a.row_begin).clamp(0, a.bank[bank].shape[0] - 1)]
        s = (rows * ehat[ctx].unsqueeze(1)).sum(-1) / a.T
        terms.append((s, held, kn, w))
    return terms, ehat


@torch.enable_grad()      # called from inside autograd.Function.forward, where grad mode is off
def _fake_forward_backward(a, ws):
    terms, ehat = _partial_terms(a)
    Z = float(a.Z)
    total = 0.0
    loss_part = []
    for s, held, kn, w in terms:
        c = kn * Z
        e = torch.exp(s)
        t = torch.cat([torch.log1p(c / e[:, :1]), torch.log1p(e[:, 1:] / c)], 1) * held
        loss_part.append(t.sum(1))
        total = total + w * t.sum() / a.mean_batch
    total.backward()
    sharded = a.row_begin != 0 or a.row_end != a.N
    lp = torch.stack([l.detach() for l in loss_part])
    if sharded:
        gv = ehat[0].grad if ehat[0].grad is not None else torch.zeros_like(ehat[0])
        ga = ehat[1].grad if ehat[1].grad is not None else torch.zeros_like(ehat[1])
        if a.grouped:       # outputs are (W, B, 128) / (W, num_keys, B) views into the buffer that is reduce-scattered
            W, Bq = a.out["grad_hat_v"].shape[:2]
            a.out["grad_hat_v"].copy_(gv.view(W, Bq, 128))
            a.out["grad_hat_a"].copy_(ga.view(W, Bq, 128))
            a.out["loss_part"].copy_(lp.view(len(a.keys), W, Bq).permute(1, 0, 2))
        else:
            a.out["grad_hat_v"].copy_(gv)
            a.out["grad_hat_a"].copy_(ga)
            a.out["loss_part"].copy_(lp)
    else:
        a.out["grad_hat_v"], a.out["grad_hat_a"], a.out["loss_part"] = ehat[0].grad, ehat[1].grad, lp
        _fake_finalize(a, ws)


@torch.enable_grad()
def _fake_finalize(a, ws):
    for ctx, name in ((0, "grad_v"), (1, "grad_a")):
        x = a.emb[ctx].detach().clone().requires_grad_(True)
        g = a.out["grad_hat_v" if ctx == 0 else "grad_hat_a"]
        (_normalize(x) * g).sum().backward()
        a.out[name].copy_(x.grad)
    lk = a.out["loss_part"].sum(1) / a.mean_batch
    a.out["loss_keys"].copy_(lk)
    a.out["loss_total"].copy_(sum(w * lk[i] for i, (_, _, _, _, w) in enumerate(a.keys)).reshape(1))


def _fake_partition_mean(a, key, out, ws):
    terms, _ = _partial_terms(SimpleNamespace(**{**a.__dict__, "keys": [a.keys[key]]}))
    s, held, kn, _ = terms[0]
    tot = (torch.exp(s[:, 1:]) * held[:, 1:]).sum()
    sharded = a.row_begin != 0 or a.row_end != a.N
    out.copy_((tot if sharded else tot / (s.shape[0] * kn)).detach().reshape(1))


def _fake_bank_update(bank_v, bank_a, emb_v, emb_a, y, mom_v, mom_a, row_begin=0, row_end=None):
    row_end = row_begin + bank_v.shape[0] if row_end is None else row_end
    emb_v, emb_a, y = _dense(emb_v, (128,)), _dense(emb_a, (128,)), _dense(y, ())
    for bank, emb, m in ((bank_v, emb_v, mom_v), (bank_a, emb_a, mom_a)):
        e = _normalize(emb)
        for i in range(y.shape[0]):
            r = int(y[i])
            if row_begin <= r < row_end:
                bank[r - row_begin] = _normalize((m * bank[r - row_begin] + (1 - m) * e[i]).view(1, -1))[0]


def _install_cpu_standins():
    from avid_cma_b200 import ops
    ops.make_nce_args = _fake_make_nce_args
    ops.nce_forward_backward = _fake_forward_backward
    ops.nce_finalize = _fake_finalize
    ops.nce_partition_mean = _fake_partition_mean
    ops.bank_update = _fake_bank_update
    ops.nce_workspace = lambda *a: torch.empty(1)
    ops.rows_l2_normalize_ = lambda x: x.copy_(_normalize(x))

    def bank_init_(bank, row_begin, seed, which):      # rows are a function of (seed, which, row), like avid_bank_init
        full = torch.randn(N, 128, generator=torch.Generator().manual_seed((seed + which) % (2 ** 63)))
        return bank.copy_(_normalize(full[row_begin:row_begin + bank.shape[0]]))
    ops.bank_init_ = bank_init_


# ---------------------------------------------------------------------------------------------- worker
def _worker(rank, world, port, backend, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    cuda = backend == "nccl"
    if cuda:
        torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank) if cuda else torch.device("cpu")
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from oracle import criterion as oc
        from oracle import synth
        if not cuda:
            _install_cpu_standins()
        from avid_cma_b200.criterions import AVID
        crit = AVID(num_data=N, embedding_dim=128, num_negatives=K, momentum=0.5, xModal_coeff=1., wModal_coeff=1.,
                    device=rank if cuda else "cpu")
        bank = crit.nce_average
        assert bank.sharded and (bank.row_begin, bank.row_end) == (rank * N // world, (rank + 1) * N // world)
        full_v, full_a = synth.bank(N, seed=5, tag="bank_v"), synth.bank(N, seed=5, tag="bank_a")
        bank.view1_mem.copy_(full_v[bank.row_begin:bank.row_end])
        bank.view2_mem.copy_(full_a[bank.row_begin:bank.row_end])
        keys = oc.avid_keys(K, 1., 1.)
        Z = -1.0
        worst = 0.0
        for step in range(2):
            embs = [synth.embeddings(B, seed=100 * step + r) for r in range(world)]
            ys = [synth.instance_ids(B, N, seed=100 * step + r) for r in range(world)]
            negs = [synth.negatives(ys[r], K, N, seed=100 * step + r) for r in range(world)]
            # unsharded oracle: every rank's loss on the full pre-update banks, Z = mean of the ranks' first-key means
            if Z <= 0:
                with torch.no_grad():
                    pms = []
                    for r in range(world):
                        sc = oc.scores(embs[r][0], embs[r][1], ys[r], full_v, full_a, negs[r], keys)
                        pms.append(oc.partition_mean(sc[keys[0].name][1]))
                    Z = float(torch.stack(pms).mean().to(torch.float32))
            ref = oc.criterion_forward_backward(embs[rank][0], embs[rank][1], ys[rank], full_v, full_a, negs[rank], keys, Z)
            for r in range(world):
                oc.bank_update(full_v, full_a, embs[r][0], embs[r][1], ys[r], 0.5)
            # sharded path
            idx = negs[rank].to(dev)
            bank.sample_negatives = lambda y_, K_, idx=idx: idx
            ev, ea = embs[rank][0].to(dev).requires_grad_(True), embs[rank][1].to(dev).requires_grad_(True)
            loss, log = crit(ev, ea, ys[rank].to(dev))
            loss.backward()

            def rel(a, b):
                a, b = a.detach().cpu().double(), b.detach().cpu().double()
                return float((a - b).norm() / b.norm().clamp_min(1e-30))
            errs = [rel(loss, ref["total"]), rel(ev.grad, ref["grad_v"]), rel(ea.grad, ref["grad_a"]),
                    abs(float(crit.criterion.avg_exp_score) - Z) / Z,
                    rel(bank.view1_mem, full_v[bank.row_begin:bank.row_end]), rel(bank.view2_mem, full_a[bank.row_begin:bank.row_end])]
            errs += [rel(log[f"Loss/{k.name}"], ref["losses"][k.name]) for k in keys]
            worst = max(worst, max(errs))
        fv, fa = bank.full_banks()
        worst = max(worst, rel(fv, full_v), rel(fa, full_a))
        # checkpoint layout: every rank takes part in the gather, the dict holds the reference's full (N,128) banks;
        # loading such a checkpoint into a sharded criterion keeps only the rows the rank owns
        from avid_cma_b200.utils import main_utils as MU
        sd = MU.reference_state_dict(crit)
        assert tuple(sd['nce_average.view1_mem'].shape) == (N, 128) and not sd['nce_average.view1_mem'].is_cuda
        worst = max(worst, rel(sd['nce_average.view1_mem'], full_v), rel(sd['nce_average.view2_mem'], full_a))
        shuffled = {k: (v.flip(0) if k.endswith('_mem') else v) for k, v in sd.items()}
        crit.load_state_dict(shuffled, strict=False)
        worst = max(worst, rel(bank.view1_mem, full_v.flip(0)[bank.row_begin:bank.row_end]))
        q.put((rank, worst, None))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _run(backend, world=2):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, worst, err in results:
        assert err is None, f"rank {rank}:\n{err}"
        assert worst < 2e-4, (rank, worst)


def test_sharded_protocol_gloo_world2():
    os.environ["AVID_SHARD_BANK"] = "1"
    try:
        _run("gloo")
    finally:
        os.environ.pop("AVID_SHARD_BANK", None)


@pytest.mark.gpu
def test_sharded_bank_nccl_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    os.environ["AVID_SHARD_BANK"] = "1"
    try:
        _run("nccl")
    finally:
        os.environ.pop("AVID_SHARD_BANK", None)


# ---------------------------------------------------------------------------------------------- in-kernel sampler across ranks
def _philox_worker(rank, world, port, q):
    """No injected negatives: both layouts draw inside the kernel from the Philox stream whose seed rank 0 broadcasts.  Each rank
    deliberately has a DIFFERENT default-generator seed (what mp.spawn workers get when only the parent was seeded)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import synth
        from avid_cma_b200.criterions import AVID
        torch.manual_seed(1000 + 17 * rank)
        n_rows, k_neg, b = 5000, 256, 8
        crits = {}
        for mode in ("sharded", "replicated"):
            os.environ["AVID_SHARD_BANK"] = "1" if mode == "sharded" else "0"
            crits[mode] = AVID(num_data=n_rows, embedding_dim=128, num_negatives=k_neg, momentum=0.5, xModal_coeff=1., wModal_coeff=1., device=rank)
        bs, br = crits["sharded"].nce_average, crits["replicated"].nce_average
        assert bs.sharded and not br.sharded and bs._seed == br._seed
        seeds = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(seeds, torch.tensor([bs._seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=dev))
        assert len({int(s) for s in seeds}) == 1, "ranks disagree on the Philox seed"
        assert torch.equal(bs.view1_mem, br.view1_mem[bs.row_begin:bs.row_end])       # per-row seeded init: same rows, no broadcast
        worst = 0.0
        for step in range(3):
            ev, ea = synth.embeddings(b, seed=10 * step + rank)
            y = synth.instance_ids(world * b, n_rows, seed=step)[rank * b:(rank + 1) * b].to(dev)
            res = {}
            for mode, crit in crits.items():
                v, a = ev.to(dev).requires_grad_(True), ea.to(dev).requires_grad_(True)
                loss, _ = crit(v, a, y)
                loss.backward()
                res[mode] = (loss.detach().double(), v.grad.double(), a.grad.double())
            worst = max(worst, float((res["sharded"][0] - res["replicated"][0]).abs() / res["replicated"][0].abs()))
            for i in (1, 2):
                worst = max(worst, float((res["sharded"][i] - res["replicated"][i]).norm() / res["replicated"][i].norm()))
            worst = max(worst, float((bs.view1_mem.double() - br.view1_mem[bs.row_begin:bs.row_end].double()).norm() / br.view1_mem.double().norm()))
            worst = max(worst, abs(float(crits["sharded"].criterion.avg_exp_score) - float(crits["replicated"].criterion.avg_exp_score)))
        q.put((rank, worst, None))
    except Exception:   # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc()))
    finally:
        os.environ.pop("AVID_SHARD_BANK", None)
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_equals_replicated_with_in_kernel_sampler_nccl_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_philox_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, worst, err in results:
        assert err is None, f"rank {rank}:\n{err}"
        assert worst < 1e-5, (rank, worst)
