#!/bin/bash
# Runs on a multi-GPU box (gpurun --gpus N -- 'bash tools/multi_gpu_check.sh N'): the product's own multi-GPU paths.
#   1. pytest: a second device driven from the same process (mandatory here: AME_EXPECT_GPUS makes a missing GPU a failure)
#   2. the CLI with --NumDevices 1 and N on a golden input: 40 log files byte-identical; per-GPU timing lines
#   3. bench.py --config shard4096 under torchrun: ONE 4096-frame batch dealt to the N ranks in blocks of 8 frames
# Everything lands in gpurun_out/r02_multi_gpu_N.log (copied to profiles/ by hand).
N=${1:-2}
O=gpurun_out/r02_multi_gpu_$N.log
export AME_EXPECT_GPUS=$N
{
  echo "== nvidia-smi -L"; nvidia-smi -L
  echo "== pytest -k second_device (AME_EXPECT_GPUS=$N)"
  timeout 300 python -m pytest tests -m gpu -q -k second_device 2>&1 | tail -3
  echo "== tools/cli_multi_gpu_check.py $N"
  timeout 600 python tools/cli_multi_gpu_check.py $N 2>&1 | tail -12; echo "cli_multi_gpu_check rc=${PIPESTATUS[0]}"
  if [ "${2:-bench}" = "bench" ]; then
    echo "== bench.py --config shard4096 --gpus $N (torchrun)"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config shard4096 --steps 1 --warmup 1 2>&1 | grep -E "^\{|rror|PARITY" | tail -3
  fi
  if [ "${3:-}" = "4k" ]; then
    echo "== bench.py --config 4k --gpus $N (torchrun)"
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --config 4k --steps 1 --warmup 1 2>&1 | grep -E "^\{|rror|PARITY" | tail -3
  fi
} > $O 2>&1
tail -30 $O
