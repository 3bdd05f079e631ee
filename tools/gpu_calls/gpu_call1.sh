#!/bin/bash
# Round-2 GPU call 1 (one B200): gather protocol on one device, new bench default, experimental kernel variants.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c1_gpus.txt 2>&1; nproc >> gpurun_out/c1_gpus.txt
echo "== gather + parity tests (default lib)"
timeout 900 python -m pytest tests/test_gpu_gather.py tests/test_gpu_parity.py tests/test_gpu_lod.py -x -q 2>&1 | tail -15 | tee gpurun_out/c1_pytest_default.log
echo "== bench default (sponza_4k), no extras"
timeout 600 python bench.py --steps 30 --no-extra > gpurun_out/c1_bench_sponza.json 2> gpurun_out/c1_bench_sponza.err; tail -c 3000 gpurun_out/c1_bench_sponza.json; tail -5 gpurun_out/c1_bench_sponza.err
for v in rcp far rcpfar; do
  echo "== variant $v: parity"
  SVX_LIB=$PWD/shocovox_b200/lib$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lod.py -x -q 2>&1 | tail -4 | tee gpurun_out/c1_pytest_$v.log
done
echo "== A/B timing"
bash tools/probe_variants.sh "- rcp far rcpfar - rcp" dot_cube_1080p cpu_render_4k colonnade_4k terrain_1024_8_1080p sponza_2048_32_4k minecraft_256_32_4k 2>&1 | tee gpurun_out/c1_variants.log
echo "== full-size parity"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -6 | tee gpurun_out/c1_pytest_fullsize.log
