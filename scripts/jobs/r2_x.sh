# InstanceNorm second stage fused into the first (ticket): tests, bench
set -x
T=r2_x
timeout 900 python -m pytest tests/test_gan_gpu.py tests/test_engine_gpu.py tests/test_iteration_parity_gpu.py -x -q > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_test.log
timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -c 300 gpurun_out/${T}_bench.err; cut -c1-220 gpurun_out/${T}_bench.json
SCDA_TIMESTAMPS=1 timeout 300 python scripts/phase_times.py > gpurun_out/${T}_phases.txt 2>&1; tail -16 gpurun_out/${T}_phases.txt
