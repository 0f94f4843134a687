#!/bin/bash
# configs[2], configs[3] and configs[4] on N GPUs of one box (default 8): one JSON line each into gpurun_out/.
#   gpurun --gpus 8 -- bash profiles/run_multi.sh 8
N=${1:-8}
mkdir -p gpurun_out
for wl in cfg2 cfg3 cfg4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --workload $wl --steps 3 --warmup 3 > gpurun_out/multi_${wl}_n${N}.json 2> gpurun_out/multi_${wl}_n${N}.err
  echo "== $wl rc=$?"; tail -c 600 gpurun_out/multi_${wl}_n${N}.json; echo; grep -E "nccl|Error|error" gpurun_out/multi_${wl}_n${N}.err | head -5
done
