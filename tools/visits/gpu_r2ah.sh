#!/bin/bash
# round 2, visit ah (2-GPU box): contract bench at N=2 with the sliced query upload in the e2e retrieval leg
TAG=r2ah
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err; echo "bench n=$N exit $?"
python - gpurun_out/scale_${TAG}_n$N.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["retrieval"]; t=d["retrieval_10M_512"]
print(d["n_gpus"], "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| 1M q/s", round(r["value"]), "e2e q/s", round(r["e2e"]["value"]), "ms", round(r["ms_per_search"],2), r["parity_spot"]["mismatch"], "| 10M q/s", round(t["value"]), "e2e", round(t["e2e"]["value"]), t["parity_spot"]["mismatch"])
PY
tail -n 3 gpurun_out/scale_${TAG}_n$N.err
