"""CMA positive mining with row-sharded banks on two ranks: every rank mines the rows it owns against ALL
candidates, the other rank's shard arrives through a broadcast, uncertified queries trigger a second -- collective -- walk over
the shard stream, and the slices are all-gathered.

CPU (`gloo`, runs everywhere): the HOST protocol of criterions/avid_cma.py::find_correspondences + ops.cma_topk with the two
kernel engines replaced by torch stand-ins (test infrastructure); rank 1's stand-in reports some queries as uncertified so
that the second walk (and rank 0's kernel-less participation in it) is exercised.  GPU (`nccl`, 2 devices): the real kernels.
The result must equal the unsharded oracle on the full banks."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, POS_K = 300 + 11, 8


class _FakeExact:
    """Torch restatement of what the fp32 engine computes: running exact top-(k+1) over the shards seen so far."""
    walks = 0

    def __init__(self, q_video, q_audio, pos_k, mode):
        self.qv, self.qa, self.pos_k, self.mode = q_video.double(), q_audio.double(), pos_k, mode
        self.val = torch.full((q_video.shape[0], 0), 0.0, dtype=torch.float64)
        self.idx = torch.zeros((q_video.shape[0], 0), dtype=torch.int64)

    def scan(self, cv, ca, begin):
        sv, sa = self.qv @ cv.double().t(), self.qa @ ca.double().t()
        sim = {"consensus": torch.minimum(sv, sa), "union": torch.maximum(sv, sa), "video": sv, "audio": sa}[self.mode]
        val = torch.cat([self.val, sim], 1)
        idx = torch.cat([self.idx, (begin + torch.arange(cv.shape[0])).expand(sim.shape[0], -1)], 1)
        k = min(self.pos_k + 1, val.shape[1])
        top = val.topk(k, dim=1).indices
        self.val, self.idx = val.gather(1, top), idx.gather(1, top)

    def finish(self):
        order = self.val.argsort(dim=1, descending=True)
        idx = self.idx.gather(1, order)[:, 1:]                  # drop the best hit (avid_cma.py:69), sort ascending (:70)
        return idx.sort(dim=1).values.to(torch.int32)


class _FakeTensorCore(_FakeExact):
    def __init__(self, q_video, q_audio, pos_k, mode, eps):
        super().__init__(q_video, q_audio, pos_k, mode)

    def finish(self):
        out = super().finish()
        rows = torch.arange(0, out.shape[0], 7) if dist.get_rank() == 1 else torch.zeros(0, dtype=torch.int64)
        out[rows] = -5                                           # garbage the re-mining pass has to repair
        return out, rows


def _worker(rank, world, port, backend, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), AVID_SHARD_BANK="1")
    cuda = backend == "nccl"
    if cuda:
        torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank) if cuda else torch.device("cpu")
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from oracle import criterion as oc
        from oracle import synth
        from avid_cma_b200 import ops
        walks = []
        if not cuda:
            ops._CmaExactEngine, ops._CmaTensorCoreEngine = _FakeExact, _FakeTensorCore
            ops.rows_l2_normalize_ = lambda x: x.copy_(torch.nn.functional.normalize(x, dim=1))
            ops.nce_workspace = lambda *a: torch.empty(1)
            ops.bank_init_ = lambda bank, row_begin, seed, which: bank.copy_(torch.nn.functional.normalize(torch.randn(
                N, 128, generator=torch.Generator().manual_seed((seed + which) % (2 ** 63)))[row_begin:row_begin + bank.shape[0]], dim=1))
        from avid_cma_b200.criterions import avid_cma
        orig = avid_cma.AVIDSimilarityPositiveExpansion._candidate_shards

        def counted(self):
            walks.append(1)
            yield from orig(self)
        avid_cma.AVIDSimilarityPositiveExpansion._candidate_shards = counted
        crit = avid_cma.AVID_CMA(num_data=N, embedding_dim=128, num_negatives=16, num_negatives_within=8, momentum=0.5,
                                 sampling_args={"type": "consensus", "pos_k": POS_K}, device=rank if cuda else "cpu")
        bank = crit.nce_average
        assert bank.sharded
        full_v, full_a = synth.bank(N, seed=31, tag="bank_v"), synth.bank(N, seed=31, tag="bank_a")
        bank.view1_mem.copy_(full_v[bank.row_begin:bank.row_end])
        bank.view2_mem.copy_(full_a[bank.row_begin:bank.row_end])
        del walks[:]
        bank.find_correspondences()
        want = oc.cma_topk(full_v.double(), full_a.double(), POS_K, "consensus")
        got = bank.positive_set.cpu()
        # index outputs must be identical; a differing row is only admissible as a proven tie at the k-th boundary (fp32 kernels vs
        # the fp64 oracle): every index of the symmetric difference has an exact similarity within 1e-6 of the boundary value
        unproven = 0
        for r in torch.nonzero((got != want).any(1)).flatten().tolist():
            sim = torch.minimum(full_v.double() @ full_v[r].double(), full_a.double() @ full_a[r].double())
            boundary = sim.sort().values[-(POS_K + 1)]
            diff = set(got[r].tolist()) ^ set(want[r].tolist())
            if cuda and all(0 <= j < N and abs(float(sim[j] - boundary)) < 1e-6 for j in diff):
                continue
            unproven += 1
        q.put((rank, unproven, len(walks), None))
    except Exception:   # noqa: BLE001
        import traceback
        q.put((rank, None, None, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _run(backend, expect_walks, world=2):
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
    for rank, mism, walks, err in results:
        assert err is None, f"rank {rank}:\n{err}"
        assert mism == 0, (rank, mism)                           # rows that differ from the oracle without being a proven tie
        if expect_walks is not None:
            assert walks == expect_walks, (rank, walks)          # BOTH ranks walked the shard stream twice (one rank had failures)


def test_sharded_cma_protocol_gloo_world2():
    _run("gloo", expect_walks=2)


@pytest.mark.gpu
def test_sharded_cma_nccl_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _run("nccl", expect_walks=1)                                 # random unit rows: every query is certified, one walk
