#!/bin/bash
# r02a: new tests first (bounded), then the whole GPU suite, then the bench (N=1) with the new tail
TAG=${1:-r02a}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q > $O/${TAG}_pytest_dp.log 2>&1; echo "pytest dp exit $?" >> $O/${TAG}_pytest_dp.log
tail -30 $O/${TAG}_pytest_dp.log
timeout 900 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -30 $O/${TAG}_pytest.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"
tail -c 3000 $O/${TAG}_bench.json; tail -20 $O/${TAG}_bench.err
