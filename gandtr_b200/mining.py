"""Training-tuple mining on the device -- SURVEY 8(f) row N3: hard-negative search and diverse-anchor selection.

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

__all__ = ["search_hard_negatives", "diverse_anchor_ranks", "select_diverse_anchors", "mark_easy_pairs", "MAX_DEPTH"]

# deepest list the device kernels rank (gdt_score_topk_exact); beyond it the walk uses the full-sort path of
# retrieval.CudaOps.exact_topk
MAX_DEPTH = 4096


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


def diverse_anchor_ranks(qpool_size, qsize, similar_exclude, similar_include, shuffle=True, generator=None):
    """The ascending rank of `most_similar` that round t of the diverse-anchor loop picks
    (cirtorch_datasets.py:84-93): `most_similar.argsort()[dissimilar_split:similar_split][choice]` is the element at rank
    dissimilar_split + choice. The splits depend only on the round number, and the random choices are drawn here with the
    same `torch.randint(n, (1,))` calls, in the same order, as the reference makes (pass a seeded generator to reproduce
    a run), so the whole schedule is known before the first kernel starts."""
    assert similar_exclude <= similar_include
    assert qsize <= qpool_size
    ranks = []
    for t in range(qsize - 1):
        valid_size = qpool_size - (t + 1)
        similar_split = max(int(valid_size * (1 - similar_exclude)), 1)
        dissimilar_split = min(int(valid_size * (1 - similar_include)), similar_split - 1)
        width = similar_split - dissimilar_split
        choice = torch.randint(width, (1,), generator=generator).item() if shuffle else width - 1
        ranks.append(dissimilar_split + choice)
    return ranks


def select_diverse_anchors(qvecs, qsize, similar_exclude, similar_include, shuffle=True, generator=None):
    """`DiverseAnchorsDataset._select_positive_pairs_db`'s greedy search (cirtorch_datasets.py:78-96).
    qvecs: [D, qpool_size] CUDA float32 (the reference's column layout). Returns (idxs, qscore_acc): the qsize selected
    pool positions (starting with 0, as the reference) and the max similarity each new anchor had to the earlier ones.
    One device loop (gdt_diverse_anchors): no per-round host sync, no growing [pool, rounds] matrix, no full sort."""
    from . import _lib
    pool = qvecs.t().contiguous()
    ranks = diverse_anchor_ranks(pool.shape[0], qsize, similar_exclude, similar_include, shuffle, generator)
    rd = torch.tensor(ranks, dtype=torch.int32).to(pool.device) if ranks else torch.empty(0, dtype=torch.int32, device=pool.device)
    picked, score = _lib.diverse_anchors(pool, rd, first=0)
    return picked.cpu().tolist(), score.cpu().tolist()


def mark_easy_pairs(qvecs, pvecs, mark_easy):
    """cirtorch_datasets.py:105-109: the `mark_easy` share of (anchor, positive) pairs with the highest similarity get the
    '-easy' label, the rest '-hard'. qvecs, pvecs: [D, qsize]."""
    qsize = qvecs.shape[1]
    sim_ord = (qvecs * pvecs).sum(0).argsort()
    easy_set = set(sim_ord[-int(mark_easy * qsize):].tolist()) if int(mark_easy * qsize) > 0 else set()
    return ["-easy" if i in easy_set else "-hard" for i in range(qsize)]
