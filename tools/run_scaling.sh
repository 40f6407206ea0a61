#!/usr/bin/env bash
# Runs bench.py at N = 1,2,4,8 for both workloads; lines go to gpurun_out/scale_<workload>.jsonl
set -u
mkdir -p gpurun_out
for wl in maxwell dc; do
  : > gpurun_out/scale_$wl.jsonl
  for n in 1 2 4 8; do
    if [ "$n" = 1 ]; then
      timeout 200 python bench.py --gpus 1 --workload $wl --steps 10 --warmup 3 --no-cpu 2>/dev/null | grep metric >> gpurun_out/scale_$wl.jsonl
    else
      timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --workload $wl --steps 10 --warmup 3 2>/dev/null | grep metric >> gpurun_out/scale_$wl.jsonl
    fi
  done
  python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/scale_$wl.jsonl")]
t1=rows[0]["ms_per_step"]
for r in rows:
    print("$wl", "N=%d"%r["n_gpus"], "%.3f ms"%r["ms_per_step"], "%.1f Gdof/s"%r["value"], "efficiency %.3f"%(t1/(r["n_gpus"]*r["ms_per_step"])), "frac/GPU %.3f"%r["roofline"]["frac"])
PY
done
