#!/bin/bash
# one GPU round: new-kernel tests (full failure text), the whole suite (summary), the default bench line, a launch list
mkdir -p gpurun_out
tag=$1
timeout 900 python -m pytest tests/test_gpu_hg.py tests/test_gpu_edges.py tests/test_gpu_boundary.py -m gpu -q --tb=short 2>&1 | grep -v "^  " | cut -c1-1800 > gpurun_out/${tag}_hg.log
timeout 900 python -m pytest tests -m gpu -q --tb=line --deselect tests/test_gpu_hg.py --deselect tests/test_gpu_edges.py --deselect tests/test_gpu_boundary.py 2>&1 | cut -c1-600 | tail -40 > gpurun_out/${tag}_all.log
timeout 120 python tools/hg_score_rng_diag.py > gpurun_out/${tag}_diag.log 2>&1
(timeout 600 python bench.py 2>gpurun_out/${tag}_bench.err | tail -2) > gpurun_out/${tag}_bench.log
tail -3 gpurun_out/${tag}_hg.log; tail -3 gpurun_out/${tag}_all.log
