# input prefetch on a copy stream + three-tap weight gradients as default: tests, bench line
set -x
T=r2_v
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_tc_gpu.py -x -q > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_test.log
timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -c 300 gpurun_out/${T}_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_v_bench.json') if l.startswith('{')][0])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
PY
