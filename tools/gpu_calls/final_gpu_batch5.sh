#!/bin/bash
# Round 2, the last GPU seconds: smoke() and the parity files that lean hardest on the host octree (full-size trees, .vox, the
# frame / edit / reload tests) on the library with the faster tree build.
mkdir -p gpurun_out
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final3.log 2>&1; tail -1 gpurun_out/r02_smoke_final3.log
timeout 115 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_vox.py tests/test_gpu_parity.py -x -q -m gpu --durations=6 2>&1 | tail -14 > gpurun_out/r02_gputest_1gpu_final3.log; cat gpurun_out/r02_gputest_1gpu_final3.log
