#!/bin/bash
# bench at N = number of visible GPUs (torchrun), with NCCL topology lines kept for the record
TAG=${1:-r02n}; N=$(nvidia-smi -L | wc -l)
O=gpurun_out; mkdir -p $O
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err; echo "bench N=$N exit $?"
grep -v "NCCL INFO" $O/${TAG}_bench_n$N.err | tail -15
grep -i "nvls\|algo\|P2P" $O/${TAG}_bench_n$N.err | sort | uniq -c | sort -rn | head -6 > $O/${TAG}_nccl_n$N.txt; cat $O/${TAG}_nccl_n$N.txt
python - <<PY
import json
d=json.loads(open("$O/${TAG}_bench_n$N.json").read())
print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "tail", d["config"]["tail"][:40])
print("tail kernels", d["roofline"]["tail_kernels_ms"])
print("nvlink", {k: v for k, v in d["roofline"].get("nvlink", {}).items() if "GBps" in k or "frac" in k})
print("variants", {k: (v.get("ms_per_step"), v.get("value"), v.get("grad_allreduce", {}).get("busbw_GBps")) if "error" not in v else v for k, v in d["variants"].items()})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_ceiling_GBps_aggregate"])
print("parity", d.get("multi_gpu_parity"), d.get("multi_gpu_parity_detail"), d.get("notes"))
PY
