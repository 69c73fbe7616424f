# third detector bucket: tests + A/B at N = 1
set -x
T=r2_ad
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_iteration_parity_gpu.py -x -q > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_test.log
for rep in 1 2; do for v in 0 1; do
SCDA_MID_BUCKET=$v timeout 300 python bench.py --steps 60 --no-cpu-baseline --no-parity-line 2> gpurun_out/${T}_$v.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('mid $v', d['ms_per_step'], d['value'])"
done; done
