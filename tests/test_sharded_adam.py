"""optim.ShardedAdam (reduce-scatter + Adam + all-gather over NVLink peer memory, csrc/shard_optim.cu) == the reference's pair
DistributedDataParallel gradient average + torch.optim.Adam (utils/main_utils.py:105-117, 250-256), on 2 ranks: parameters after
several steps, torch.optim.Adam's state_dict layout (gathered moments), resume from that state_dict, LocalGradients' key prefix."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(64, 3, 3, 7, 7), (64,), (130, 7), (5,), (512, 512), (128, 64, 1, 3, 3), (3,)]      # ragged sizes: not multiples of 4, fewer elements than ranks


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from avid_cma_b200 import optim
        assert optim.ShardedAdam.available()
        g = torch.Generator().manual_seed(1)                       # same initial parameters on every rank
        init = [torch.randn(s, generator=g) for s in SHAPES]
        mine = [torch.nn.Parameter(t.clone().to(dev) + (0.5 if rank else 0.0)) for t in init]      # rank 1 starts perturbed: the constructor must sync
        ref = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
        kw = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
        opt = optim.ShardedAdam(mine, **kw)
        ref_opt = torch.optim.Adam(ref, **kw)
        worst = 0.0

        def grads(step):
            gg = torch.Generator().manual_seed(100 * step + rank)
            return [torch.randn(s, generator=gg).to(dev) for s in SHAPES]

        def one_step(step, o, ps, r_o, r_ps):
            gs = grads(step)
            for p, x in zip(ps, gs):
                p.grad = x.clone()
            if step == 2:
                ps[3].grad = None                                  # a parameter without a gradient counts as zero on this rank
                gs[3] = torch.zeros_like(gs[3])
            o.step()
            for p, x in zip(r_ps, gs):                             # DDP: average over the ranks, then Adam
                avg = x.clone()
                dist.all_reduce(avg)
                p.grad = avg / world
            r_o.step()

        detail = []
        for step in range(1, 5):
            one_step(step, opt, mine, ref_opt, ref)
            errs = [float((a.detach() - b.detach()).abs().max() / b.detach().abs().max()) for a, b in zip(mine, ref)]
            detail.append(('step', step, ['%.1e' % e for e in errs]))
            worst = max(worst, max(errs))
        # state_dict in torch.optim.Adam's layout: load it into torch's Adam and into a fresh ShardedAdam, continue, compare
        sd = opt.state_dict()
        assert set(sd) == {'state', 'param_groups'} and len(sd['state']) == len(SHAPES)
        assert all(tuple(sd['state'][i]['exp_avg'].shape) == tuple(s) for i, s in enumerate(SHAPES))
        rsd = ref_opt.state_dict()
        for i in range(len(SHAPES)):
            for k in ('exp_avg', 'exp_avg_sq'):
                d = float((sd['state'][i][k] - rsd['state'][i][k]).abs().max() / rsd['state'][i][k].abs().max().clamp_min(1e-30))
                detail.append((k, i, '%.1e' % d))
                worst = max(worst, d)
            assert int(sd['state'][i]['step']) == int(rsd['state'][i]['step']) == 4
        mine2 = [torch.nn.Parameter(p.detach().clone()) for p in mine]
        opt2 = optim.ShardedAdam(mine2, **kw)
        opt2.load_state_dict(sd)
        for step in range(5, 7):
            one_step(step, opt2, mine2, ref_opt, ref)
            errs = [float((a.detach() - b.detach()).abs().max() / b.detach().abs().max()) for a, b in zip(mine2, ref)]
            detail.append(('resumed step', step, ['%.1e' % e for e in errs]))
            worst = max(worst, max(errs))
        # the wrapper that stands in for DistributedDataParallel
        lin = torch.nn.Linear(4, 4).to(dev)
        wrapped = optim.LocalGradients(lin)
        assert list(wrapped.state_dict().keys()) == ['module.weight', 'module.bias'] and wrapped.module is lin
        q.put((rank, worst, None if worst < 2e-6 else 'mismatch: %r' % (detail,)))
    except Exception:   # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_adam_equals_ddp_average_plus_adam_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, worst, err in results:
        assert err is None, f"rank {rank}:\n{err}"
        assert worst < 2e-6, (rank, worst)
