#!/bin/bash
# scaling curves on one box: weak (driver metric), strong (fixed ensemble), cfg4 at the full rank count
NMAX=${1:-8}
port=29600
run() { # N extra-args tag
  local n=$1; shift; local tag=$1; shift
  port=$((port+1))
  if [ "$n" = "1" ]; then python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu "$@" > gpurun_out/r2_scale_${tag}_n$n.json 2> gpurun_out/r2_scale_${tag}_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 100 --warmup 5 --no-cpu "$@" > gpurun_out/r2_scale_${tag}_n$n.json 2> gpurun_out/r2_scale_${tag}_n$n.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_scale_${tag}_n$n.json").read().strip().splitlines()[-1])
    print("${tag} N=$n", "ms/step", round(d["ms_per_step"],4), "value", "%.3e"%d["value"], "e2e_ms", round(d["e2e"]["ms_per_step"],4), "merge_ms", round(d["roofline"]["merge_ms_per_step"],4), "scan_ms", round(d["roofline"]["kernel_ms_per_step"],4), "parity", d["parity_checked"])
except Exception as e:
    print("${tag} N=$n FAILED", e); print(open("gpurun_out/r2_scale_${tag}_n$n.err").read()[-800:])
PY
}
for n in 1 2 4 8; do [ $n -le $NMAX ] && run $n weak; done
for n in 2 4 8; do [ $n -le $NMAX ] && run $n strong --scaling strong; done
[ $NMAX -ge 8 ] && run 8 cfg4 --config cfg4
