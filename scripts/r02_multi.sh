#!/bin/bash
# Multi-GPU measurements: gpurun --gpus N --timeout 1500 -- 'bash scripts/r02_multi.sh N [tests] [c5]'
set -u
N=$1; shift
O=gpurun_out/r02f
mkdir -p $O
what="$*"
PORT=29512
if [[ " $what " == *" tests "* ]]; then
  timeout 900 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_parity.py tests/test_gpu_gradients.py tests/test_device_shell.py -m gpu -q > $O/gpu_tests_n$N.log 2>&1
  tail -2 $O/gpu_tests_n$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
  bench.py --gpus $N > $O/bench_default_n$N.json 2> $O/bench_default_n$N.err
echo "bench N=$N rc=$?"; tail -c 400 $O/bench_default_n$N.err
if [[ " $what " == *" c5 "* ]]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --gpus $N --workload c5 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err
  echo "bench c5 N=$N rc=$?"
fi
python - <<PY
import json
for f in ("$O/bench_default_n$N.json", "$O/bench_c5_n$N.json"):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        g = d.get("gram")
        print(f, d["n_gpus"], round(d["ms_per_step"], 4), d["value"], "gram", None if g is None else (round(g["ms_per_step"], 2), g["value"], g.get("stages_ms_rank0")))
    except Exception as e:
        print(f, "n/a", e)
PY
