#!/bin/bash
# scaling visit on an N-GPU box: contract bench at every N in {1,2,4,8} <= available, plus the 10M x 512 database config
NMAX=${1:-8}
TAG=${2:-rX}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for N in 1 2 4 8; do
  if [ $N -le $NMAX ]; then
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
    fi
    echo "bench n=$N exit $?" >> gpurun_out/summary.txt
  fi
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $NMAX --steps 5 --warmup 3 --db-rows 10000000 --db-dim 512 > gpurun_out/scale_${TAG}_10M512_n$NMAX.json 2> gpurun_out/scale_${TAG}_10M512_n$NMAX.err; echo "bench 10M n=$NMAX exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --db-rows 10000000 --db-dim 512 > gpurun_out/scale_${TAG}_10M512_n1.json 2> gpurun_out/scale_${TAG}_10M512_n1.err; echo "bench 10M n=1 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/scale_${TAG}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d.get("retrieval",{})
    print(d["n_gpus"], "img/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| retrieval q/s", round(r.get("value",0)), "ms", round(r.get("ms_per_search",0),2), "frac", round(r.get("roofline",{}).get("frac",0),3), r.get("status"))
except Exception as e:
    print("unparsable:", e); print(open(sys.argv[1]).read()[:500])
PY
done
tail -n 5 gpurun_out/scale_${TAG}_*.err
