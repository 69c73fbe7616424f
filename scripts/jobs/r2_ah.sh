# fresh small gradients are not zeroed twice: tests + step time
set -x
T=r2_ah
timeout 100 python -m pytest tests/test_engine_gpu.py tests/test_disc_gpu.py tests/test_gan_gpu.py tests/test_iteration_parity_gpu.py tests/test_tc_detector_gpu.py -x -q > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_test.log
timeout 60 python bench.py --steps 50 --no-cpu-baseline --no-parity-line 2> gpurun_out/${T}.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(d['ms_per_step'], d['value'], d['gpu_launches'])"
