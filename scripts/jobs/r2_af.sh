# fused single-image prediction kernels: tests + config 2
set -x
T=r2_af
timeout 600 python -m pytest tests/test_eval_gpu.py tests/test_functions_gpu.py tests/test_driver_gpu.py tests/test_model_gpu.py -x -q > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -15 gpurun_out/${T}_test.log
timeout 300 python bench.py --config 2 > gpurun_out/${T}_config2.json 2> gpurun_out/${T}_config2.err; echo rc=$?; cut -c1-400 gpurun_out/${T}_config2.json; tail -3 gpurun_out/${T}_config2.err
