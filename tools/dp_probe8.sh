#!/bin/bash
# 8-GPU probe: weak (64 / GPU) and strong (global 64) scaling points, graph vs eager, bucket sizes
N=${1:-8}
run() { local label=$1; shift; local envs=(); while [[ $1 != "--" ]]; do envs+=("$1"); shift; done; shift
  local out=$(env "${envs[@]}" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-bandwidth --no-extras --no-inference "$@" 2>/dev/null | tail -1)
  echo "$out" > gpurun_out/dp8_$(echo $label | tr ' ' '_').json
  echo "$label: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "samples/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), d.get("dp_check",{}).get("status"), d["clocks"]["sm_mhz"])' 2>&1 | tail -1)"
}
run "weak graph 4MB" X=1 --
run "weak graph single" POSEB200_BUCKET_BYTES=268435456 --
run "strong graph single" POSEB200_BUCKET_BYTES=268435456 -- --batch-per-gpu $((64 / N))
run "strong graph 4MB" X=1 -- --batch-per-gpu $((64 / N))
run "strong eager 4MB" X=1 -- --batch-per-gpu $((64 / N)) --no-graph
