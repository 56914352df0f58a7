"""Row-sharded retrieval over NCCL on 2 GPUs (skipped on single-GPU boxes): per-shard tcgen05 filter, histogram exchange
(all-reduce), finalize against the GLOBAL threshold, all_gather + merge -- identical to the single-shard oracle ranking."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import retrieval_np as R
from tests.util import unit_rows

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from gandtr_b200.retrieval import ShardedIndex, evaluate_map
        rs = np.random.RandomState(7)
        nq, ndb, d, k = 150, 60001, 256, 100
        db = unit_rows(rs, ndb, d)
        src = rs.randint(0, ndb, nq)
        q = db[src] + 0.4 * rs.normal(0, 1, (nq, d)).astype(np.float32) / np.sqrt(d)
        q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
        index = ShardedIndex.from_full(torch.from_numpy(db).cuda())
        qd = torch.from_numpy(q).cuda() if rank == 0 else torch.zeros((nq, d), device="cuda")
        s, i = index.search(qd, k, broadcast=True)
        os_, oi = R.topk(R.scores_exact(q, db), k)
        assert np.array_equal(i.cpu().numpy(), oi), "rank %d: merged index lists differ from the oracle" % rank
        assert np.abs(s.cpu().numpy() - os_).max() < 1e-6
        st = index.shard.last_status
        assert st is not None and st[0] == 0
        # the exchange makes each shard re-score only its share of the global top k (+ margin), not a full local top k
        assert st[1] < k + 40, st
        gnd = [{"ok": np.array([src[j]]), "junk": np.array([(src[j] + 1) % ndb])} for j in range(nq)]
        m, aps, _, _ = evaluate_map(index, qd, gnd)
        mo, apo, _, _ = R.compute_map(R.full_ranks(R.scores_exact(q, db)), gnd)
        assert abs(m - mo) < 1e-12 and np.array_equal(aps, apo)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_search_two_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
