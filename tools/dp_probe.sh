#!/bin/bash
# data-parallel step time under different bucket sizes / NCCL CTA limits / launch modes.  usage: tools/dp_probe.sh N
N=${1:-2}
run() { # label, env..., -- extra args
  local label=$1; shift
  local envs=(); while [[ $1 != "--" ]]; do envs+=("$1"); shift; done; shift
  local out=$(env "${envs[@]}" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-bandwidth --no-extras --no-inference "$@" 2>/dev/null | tail -1)
  echo "$label: $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]), "samples/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"]), d.get("dp_check",{}).get("status"))' 2>&1 | tail -1)"
}
run "graph  4MB buckets        " X=1 --
run "graph  single bucket      " POSEB200_BUCKET_BYTES=268435456 --
run "graph  1MB buckets        " POSEB200_BUCKET_BYTES=1048576 --
run "graph  4MB NCCL_MAX_CTAS=2" NCCL_MAX_CTAS=2 --
run "graph  4MB NCCL_MAX_CTAS=8" NCCL_MAX_CTAS=8 --
run "eager  4MB buckets        " X=1 -- --no-graph
run "eager  single bucket      " POSEB200_BUCKET_BYTES=268435456 -- --no-graph
