# warp-per-channel second stage of InstanceNorm / first-layer weight gradient: tests, bench
set -x
T=r2_ab
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_disc_gpu.py tests/test_engine_gpu.py tests/test_iteration_parity_gpu.py -x -q > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_test.log
timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; cut -c1-220 gpurun_out/${T}_bench.json
