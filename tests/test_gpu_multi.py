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


def _score_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from gandtr_b200 import hub
        from gandtr_b200.extract import extract_descriptors
        from gandtr_b200.score import CirDatasetAp
        from tests.util import synth_image
        torch.backends.cudnn.allow_tf32 = False
        torch.manual_seed(0)
        net = hub.gem_vgg16_hedngan(pretrained=False, device="cuda:%d" % rank)
        images = [synth_image(1000 + i, 64, 80, "smooth" if i % 4 else "noise") for i in range(21)]
        qimages = [np.clip(images[j].astype(np.int16) + 8, 0, 255).astype(np.uint8) for j in range(4)]
        gnd = [{"ok": np.array([j]), "junk": np.array([(j + 5) % 21])} for j in range(4)]
        data = net.network_params.runtime["data"]
        score = CirDatasetAp({"image_size": None, "dataset": {"name": "syn", "images": images, "qimages": qimages,
                                                              "bbxs": [None] * 4, "gnd": gnd},
                              "transforms": data.get("transforms", data.get("augmentations")), "mean_std": data["mean_std"]})
        avg = score(net, "cuda", lambda *a: None)                       # database extraction and index sharded over 2 ranks
        db = extract_descriptors(net, images, None, net.transform).cpu().numpy()
        q = extract_descriptors(net, qimages, None, net.transform).cpu().numpy()
        m, _, _, _ = R.compute_map(R.full_ranks(R.scores_exact(q, db)), gnd)
        assert abs(avg["map"] - m) < 1e-9, (avg, m)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_score_object_sharded_over_two_gpus():
    """CirDatasetAp with the database extraction and the index sharded over 2 ranks equals the single-process oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_score_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)
