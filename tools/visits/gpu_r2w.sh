#!/bin/bash
# round 2, visit w (4-GPU box): contract bench at N=2 and N=4 (what the driver's scaling run launches)
TAG=r2w
mkdir -p gpurun_out; rm -f gpurun_out/summary_$TAG.txt
for N in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err; echo "bench n=$N exit $?" >> gpurun_out/summary_$TAG.txt
done
cat gpurun_out/summary_$TAG.txt
for N in 2 4; do python - gpurun_out/scale_${TAG}_n$N.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get("retrieval",{}); t=d.get("retrieval_10M_512",{})
    print(d["n_gpus"], "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| 1M q/s", round(r.get("value",0)), "ms", round(r.get("ms_per_search",0),2), "frac", round(r.get("roofline",{}).get("frac",0),3), r.get("status"), r.get("parity_spot",{}).get("mismatch"), "| 10M q/s", round(t.get("value",0)), "ms", round(t.get("ms_per_search",0),2), t.get("parity_spot",{}).get("mismatch"))
except Exception as e:
    print("unparsable:", e); print(open(sys.argv[1]).read()[:300])
PY
tail -n 3 gpurun_out/scale_${TAG}_n$N.err; done
