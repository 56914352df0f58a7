"""ncu target: the fused score + top-k kernel at d = 512 (BASELINE config 5 shape, reduced)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from gandtr_b200 import _lib
nq, ndb, d = 4096, 524288, 512
db = torch.randn((ndb, d), device="cuda"); db /= db.norm(dim=1, keepdim=True)
q = torch.randn((nq, d), device="cuda"); q /= q.norm(dim=1, keepdim=True)
shadow, stats = _lib.db_prepare(db)
for _ in range(3):
    s, i, st = _lib.score_topk(q, db, shadow, stats, 100)
torch.cuda.synchronize()
print("status", st.cpu().tolist())
