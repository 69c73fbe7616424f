# three-tap weight-gradient form: per-layer times and step time, A/B
set -x
T=r2_u
SCDA_WGRAD3=1 timeout 300 python scripts/convbench.py > gpurun_out/${T}_convbench_wg3.jsonl 2> gpurun_out/${T}_convbench_wg3.err; echo rc=$?; tail -2 gpurun_out/${T}_convbench_wg3.err
SCDA_WGRAD3=1 timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench_wg3.json 2> gpurun_out/${T}_bench_wg3.err; echo rc=$?; cut -c1-200 gpurun_out/${T}_bench_wg3.json
SCDA_WGRAD3=0 timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench_wg1.json 2> gpurun_out/${T}_bench_wg1.err; echo rc=$?; cut -c1-200 gpurun_out/${T}_bench_wg1.json
