#!/bin/bash
# Round 2, last call (one B200, ~7 GPU-minutes left): the whole GPU suite, smoke() and the default bench line on the final library.
mkdir -p gpurun_out
timeout 280 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_gputest_1gpu_final2.log; cat gpurun_out/r02_gputest_1gpu_final2.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final2.log 2>&1; tail -2 gpurun_out/r02_smoke_final2.log
timeout 100 python bench.py --no-extra > gpurun_out/r02_bench_n1_final2.json 2> gpurun_out/r02_bench_n1_final2.err; head -c 400 gpurun_out/r02_bench_n1_final2.json; tail -c 300 gpurun_out/r02_bench_n1_final2.err
