# pair form of the halo kernel + direct conv1_1: parity tests, tile-plan sweep, step time
set -x
T=r2_p
timeout 600 python -m pytest tests/test_conv_halo_gpu.py -x -q > gpurun_out/${T}_halotest.log 2>&1; echo rc=$?; tail -15 gpurun_out/${T}_halotest.log
timeout 600 python scripts/halobench.py > gpurun_out/${T}_halobench.jsonl 2> gpurun_out/${T}_halobench.err; echo rc=$?; tail -3 gpurun_out/${T}_halobench.err
timeout 900 python -m pytest tests/test_tc_detector_gpu.py tests/test_engine_gpu.py tests/test_iteration_parity_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/${T}_dettest.log 2>&1; echo rc=$?; tail -8 gpurun_out/${T}_dettest.log
timeout 600 python bench.py --steps 30 --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -c 600 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
SCDA_TIMESTAMPS=1 timeout 300 python scripts/phase_times.py > gpurun_out/${T}_phases.txt 2>&1; tail -25 gpurun_out/${T}_phases.txt
