set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -c 600 gpurun_out/bench_r1f.err
cut -c1-400 gpurun_out/bench_r1f.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f_step_eager.csv python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/ncu_bench_f.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/conv_r1f -f python scripts/prof_conv.py > gpurun_out/prof_conv_f.log 2>&1
python scripts/convbench.py > gpurun_out/convbench_r1c.jsonl 2>&1
