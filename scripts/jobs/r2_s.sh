# k-means beside the RCNN head, GAN chain forked at the crops, masks on a helper stream: tests + step time
set -x
T=r2_s
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_iteration_parity_gpu.py tests/test_model_gpu.py tests/test_tc_detector_gpu.py tests/test_loss_ops_gpu.py tests/test_kmeans_gpu.py tests/test_driver_gpu.py -x -q > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -5 gpurun_out/${T}_test.log
timeout 600 python bench.py --steps 50 --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -c 600 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
SCDA_TIMESTAMPS=1 timeout 300 python scripts/phase_times.py > gpurun_out/${T}_phases.txt 2>&1; tail -16 gpurun_out/${T}_phases.txt
