#!/bin/bash
TAG=${1:-r02am}; O=gpurun_out; mkdir -p $O
ls oracle/_ref | head -3; ls /root/reference 2>&1 | head -1
timeout 600 python -m pytest tests/test_gpu_vs_reference.py -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -15 $O/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; tail -c 1500 $O/${TAG}_bench_reference.json
