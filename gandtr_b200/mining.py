"""Hard-negative search for training-tuple mining -- SURVEY 8(f) row N3.

Replaces `TuplesDataset._search_hard_negatives` (mdir/external/cirtorch/datasets/traindataset.py:246-279), which sorts
the full pool x query score matrix (`torch.sort` of e.g. 22 000 x 2 000) and then walks each ranked column on the host.
Here the ranking comes from the fused score + top-k kernel (K3): only the first `depth` ranks per query are ever
produced, and the cluster-exclusion walk (never the query's cluster, at most one image per cluster) runs over that short
list; the depth doubles for the rare queries whose list is exhausted, so the result is identical to the reference's walk
over the full ranking (ties in score aside: the library's total order is score desc, index asc).
"""
import numpy as np
import torch

from .retrieval import ShardedIndex

__all__ = ["search_hard_negatives"]


def search_hard_negatives(qvecs, poolvecs, qclusters, poolclusters, nnum, idxs2images=None, depth=None):
    """qvecs: [D, nq], poolvecs: [D, npool] (the reference's column-vector layout) CUDA float32.
    qclusters [nq], poolclusters [npool]: cluster id of every query / pool image. Returns (nidxs, stats):
    nidxs[q] = list of `nnum` pool indices (mapped through `idxs2images` when given), stats as the reference's
    {"average_negative_distance": [...]} (sqrt(sum((q - x + 1e-6)^2)) per selected negative)."""
    nq, npool = qvecs.shape[1], poolvecs.shape[1]
    q = qvecs.t().contiguous()
    pool = poolvecs.t().contiguous()
    index = ShardedIndex(pool)
    qclusters = np.asarray(qclusters)
    poolclusters = np.asarray(poolclusters)
    depth = int(depth or max(4 * nnum + 16, 64))
    nidxs, dists = [None] * nq, [None] * nq
    todo = np.arange(nq)
    while len(todo):
        k = min(depth, npool)
        _, idx = index.search(q[torch.as_tensor(todo, device=q.device)].contiguous(), k)
        idx = idx.cpu().numpy()
        again = []
        for row, qi in enumerate(todo):
            clusters, chosen = {qclusters[qi]}, []
            for r in range(k):
                cand = int(idx[row, r])
                if cand < 0:
                    break
                c = poolclusters[cand]
                if c not in clusters:
                    chosen.append(cand)
                    clusters.add(c)
                    if len(chosen) == nnum:
                        break
            if len(chosen) < nnum and k < npool:
                again.append(qi)
                continue
            nidxs[qi] = chosen
        if again and k >= npool:
            break
        todo, depth = np.asarray(again, dtype=np.int64), depth * 2
    sel = torch.as_tensor([c for row in nidxs for c in row], device=q.device, dtype=torch.long)
    owner = torch.as_tensor([qi for qi, row in enumerate(nidxs) for _ in row], device=q.device, dtype=torch.long)
    nd = torch.pow(q[owner] - pool[sel] + 1e-6, 2).sum(dim=1).sqrt().cpu().tolist() if len(sel) else []
    if idxs2images is not None:
        nidxs = [[idxs2images[c] for c in row] for row in nidxs]
    return nidxs, {"average_negative_distance": nd}
