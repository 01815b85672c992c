#!/bin/bash
# Final measurements of round 2 on one B200 (gpurun --timeout 2400 -- 'bash scripts/r02_final.sh [stage ...]'); everything
# lands under gpurun_out/r02f/.  Stages: tests bench others ref launches ncu
set -u
O=gpurun_out/r02f
mkdir -p $O
what="${*:-tests bench others ref launches ncu}"
has() { [[ " $what " == *" $1 "* ]]; }

if has tests; then
  timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
fi
if has bench; then
  timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
fi
if has others; then
  for w in c4 c5 c5c; do
    timeout 900 python bench.py --workload $w > $O/bench_$w.json 2> $O/bench_$w.err; echo "bench $w rc=$?"
  done
fi
if has ref; then
  timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sustained > $O/launches_bench.log 2>&1
  echo "launch list rc=$?"
fi
cap() {  # cap <name> <kernel regex> <skip> <target args...>
  local name=$1 k=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $skip -c 1 -f -o $O/$name \
    python scripts/profile_target.py "$@" > $O/$name.log 2>&1
  echo "ncu $name rc=$?"
}
if has ncu; then
  cap c2_expval_cov        'expval_cov_kernel'            2 c2
  cap c4_circuit_columns   'circuit_columns_kernel'       1 c4cols 16384
  cap c4_circuit_panel     'circuit_panel_kernel'         1 c4 16384
  cap pf128_static         'pfaffian_static128_kernel'    2 pfcov 128
  cap pf256_panel          'pfaffian_panel_kernel'        2 pfcov 256
fi
ls -la $O
