#!/bin/bash
# r02b: GPU suite + bench at N = number of visible GPUs
TAG=${1:-r02b}; N=$(nvidia-smi -L | wc -l)
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -12 $O/${TAG}_pytest.log
if [ "$N" -gt 1 ]; then
  NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "bench N=$N exit $?"
  tail -c 4000 $O/${TAG}_bench_n$N.json; grep -v "NCCL INFO" $O/${TAG}_bench_n$N.err | tail -20; grep -i "nvls\|Using network\|via P2P\|algo" $O/${TAG}_bench_n$N.err | head -8
else
  timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
  tail -c 2500 $O/${TAG}_bench.json; tail -5 $O/${TAG}_bench.err
fi
