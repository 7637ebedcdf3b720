#!/usr/bin/env bash
# BASELINE configs[2..4] on one 8 x B200 box (one gpurun --gpus 8 call):
#   [2] Qwen3-8B-shaped 3-bit asym g128: layers sharded over N ranks + one block token-sharded with the NCCL all-reduce
#   [3] Qwen3-8B-shaped 2-bit asym g128, eps sweep (the retained rank k moves with eps)
#   [4] Llama-3-70B-shaped solver sweep, n = 8192 and n = 28672, one replica per GPU
set -u
cd "$(dirname "$0")/.."
N="${1:-8}"
out=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29533 "$@"; }
run bench.py --gpus "$N" --steps 3 --warmup 2 --bits 3 --sym 0 --e2e-steps 1 > $out/r02_cfg2_w3a_n$N.json 2> $out/r02_cfg2_w3a_n$N.err
tail -1 $out/r02_cfg2_w3a_n$N.err
for eps in 1e-7 1e-5 1e-2; do
  run bench.py --gpus "$N" --steps 2 --warmup 2 --bits 2 --sym 0 --eps $eps --no-extras --e2e-steps 0 > $out/r02_cfg3_w2a_eps${eps}_n$N.json 2> $out/r02_cfg3_w2a_eps${eps}_n$N.err
  tail -1 $out/r02_cfg3_w2a_eps${eps}_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29534 scripts/llama70b_sweep.py 8192 28672 > $out/r02_llama70b_n$N.jsonl 2> $out/r02_llama70b_n$N.err
cat $out/r02_llama70b_n$N.jsonl
