#!/bin/bash
# usage (on a multi-GPU box): bash tools/run_scaling.sh "8" "" "8" "--no-pipeline" "8" "--shard channels" "4" "" ...
# pairs of (N, extra bench.py flags); writes gpurun_out/scale_<idx>_G<N>.json and prints a one-line summary each
port=29510
idx=0
while [ $# -ge 2 ]; do
    N=$1; FLAGS=$2; shift 2
    port=$((port + 1)); idx=$((idx + 1))
    out=gpurun_out/scale_${idx}_G${N}.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$port" \
        bench.py --gpus "$N" --steps 50 --warmup 3 $FLAGS 2> gpurun_out/scale_${idx}_G${N}.err | tail -1 > "$out"
    python - "$out" "$FLAGS" <<'PY' || tail -5 gpurun_out/scale_${idx}_G${N}.err
import json, sys
d = json.load(open(sys.argv[1]))
r = d.get("roofline", {})
print(d["n_gpus"], repr(sys.argv[2]), "value", round(d["value"]), "e2e", round(d.get("e2e", {}).get("value", 0)),
      "ms/step", round(d["ms_per_step"], 3), r.get("phases_ms_per_step"), d["clocks"]["reasons"])
PY
done
