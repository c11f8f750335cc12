#!/bin/bash
# one GPU round (round 2): the whole -m gpu suite, smoke(), the default bench line, launch list + ncu --set full of the wide kernels
mkdir -p gpurun_out
tag=$1
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^  " | cut -c1-1500 | tail -60 > gpurun_out/${tag}_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1)
(timeout 900 python bench.py --steps 20 --warmup 3 2>gpurun_out/${tag}_bench.err | tail -2) > gpurun_out/${tag}_bench.log
tail -3 gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_smoke.log
