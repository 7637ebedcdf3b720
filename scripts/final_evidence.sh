#!/usr/bin/env bash
# One gpurun call that produces the artefacts profiles/README.md lists for the end of a round:
#   the default bench line, the full GPU test log, the ncu launch list of a short bench run (plain run first),
#   and `ncu --set full` summaries of the kernels that changed (pivoted Cholesky panel, error-metric GEMM).
set -u
cd "$(dirname "$0")/.."
out=gpurun_out
tag="${1:-r02_final}"
python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
tail -3 $out/${tag}_bench_n1.err
python -m pytest tests -m gpu -q > $out/${tag}_gpu_tests.log 2>&1; tail -2 $out/${tag}_gpu_tests.log
short="bench.py --steps 1 --warmup 1 --no-extras --e2e-steps 0 --no-cpu-baseline"
if TQ_BENCH_NO_SMI=1 python $short > $out/${tag}_short.json 2> $out/${tag}_short.err; then
  TQ_BENCH_NO_SMI=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
    --log-file $out/${tag}_launches_bench.csv python $short > $out/${tag}_ncu_bench.log 2>&1
  echo "ncu launch list rc=$?"
fi
if python scripts/ncu_solver_target.py 12288 > $out/${tag}_solver_target.log 2>&1; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pchol_panel_kernel -s 40 -c 2 \
    -o $out/${tag}_ncu_pchol python scripts/ncu_solver_target.py 12288 > $out/${tag}_ncu_pchol.log 2>&1
  echo "ncu pchol rc=$?"
fi
if python scripts/loop_probe.py > $out/${tag}_loop_probe.log 2>&1; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:metric_tc_kernel -c 1 \
    -o $out/${tag}_ncu_metric python scripts/loop_probe.py > $out/${tag}_ncu_metric.log 2>&1
  echo "ncu metric rc=$?"
fi
cat $out/${tag}_loop_probe.log
