#!/bin/bash
# One gpurun call that collects the round's evidence: GPU tests, the default bench line, the ncu launch list of the same
# command and one `ncu --set full` capture per kernel the north_star names.  Everything lands under gpurun_out/r02/.
#   gpurun --timeout 1500 -- 'bash scripts/r02_evidence.sh [tests] [bench] [launches] [ncu]'
set -u
O=gpurun_out/r02
mkdir -p $O
what="${*:-tests bench launches ncu}"
has() { [[ " $what " == *" $1 "* ]]; }

if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1
  tail -3 $O/gpu_tests.log
fi
if has bench; then
  timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
  echo "bench rc=$?"; head -c 600 $O/bench_default.json; echo
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv \
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
  cap c3_gram_tiles        'gram_tiles'                   1 gram 2048
  cap c4_circuit_columns   'circuit_columns'              1 c4 16384
  cap c4_cov_basis         'cov_basis_kernel'             1 c4 16384
  cap c4_probs_tree        'probs_tree_kernel'            1 c4 16384
  cap pf128_block_tiles    'pfaffian_block_tiles_kernel'  2 pf 128
  cap pf256_block_tiles    'pfaffian_block'               2 pf 256
  cap matmul32_tma         'sptm_matmul_tma_kernel'       2 matmul 32
  cap matmul64_tma         'sptm_matmul_tma_kernel'       2 matmul 64
  cap matmul128_dmma       'bgemm_dmma_kernel'            2 matmul 128
  cap matmul256_dmma       'bgemm_dmma_kernel'            2 matmul 256
  cap c5_gram_blocks4      'gram_blocks4_kernel'          1 blocks 32768
fi
ls -la $O
