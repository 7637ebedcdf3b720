#!/usr/bin/env bash
# Round-2 opener for the experimental two-stage tridiagonal reduction: the gated GPU parity tests,
# then tq_eigh one-stage vs two-stage for every variant that was written after the first (and only) GPU measurement.
# The library reads its switches once per process, hence one process per variant.
#   gpurun --timeout 600 -- 'bash scripts/two_stage_variants.sh > gpurun_out/two_stage_variants.log 2>&1'
set -u
cd "$(dirname "$0")/.."
sizes="${*:-4096 12288}"
echo "== gated parity tests (default variant)"
TQ_TEST_TWO_STAGE=1 timeout 300 python -m pytest tests/test_gpu_two_stage.py -x -q 2>&1 | tail -5
for v in "" "TQ_CHASE_HELPER=1" "TQ_CHASE_LATE=1" "TQ_CHASE_HELPER=1 TQ_CHASE_LATE=1" "TQ_SY2SB_GEMM=1" \
         "TQ_SY2SB_LOOKAHEAD=1" "TQ_SY2SB_LOOKAHEAD=1 TQ_CHASE_HELPER=1 TQ_CHASE_LATE=1"; do
  echo "== variant: ${v:-default}"
  # shellcheck disable=SC2086
  env $v TQ_TRACE=1 timeout 120 python scripts/two_stage_probe.py $sizes 2>&1 | grep -E "sy2sb|sb2st|apply_q|^\{" | tail -24
done
echo "== parity tests once more with every new variant switched on"
TQ_SY2SB_LOOKAHEAD=1 TQ_CHASE_HELPER=1 TQ_CHASE_LATE=1 TQ_TEST_TWO_STAGE=1 timeout 300 python -m pytest tests/test_gpu_two_stage.py -x -q 2>&1 | tail -5
echo "== whole solver (tq_spectral_solve) at n = 12288: one-stage, then two-stage with every new variant"
timeout 120 python scripts/solver_sweep.py 12288 2>&1 | tail -1
TQ_EIGH_TWO_STAGE=1 TQ_SY2SB_LOOKAHEAD=1 TQ_CHASE_HELPER=1 TQ_CHASE_LATE=1 timeout 120 python scripts/solver_sweep.py 12288 2>&1 | tail -1
