# unsplit one-tap weight gradient for the short-reduction layers: tests, per-layer time, A/B of the iteration
set -x
T=r2_ag
timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_tc_detector_gpu.py -x -q -k "wgrad or backbone or rpn or detector" > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_test.log
for v in 1 0 1 0; do
SCDA_WGRAD_SHORT_K=$v timeout 200 python bench.py --steps 60 --no-cpu-baseline --no-parity-line 2> gpurun_out/${T}_$v.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('short_k $v', d['ms_per_step'], d['value'])"
done
