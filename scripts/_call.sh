set -x
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -c 300 gpurun_out/bench_r1i.err
cut -c1-300 gpurun_out/bench_r1i.json
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/conv_r1i -f python scripts/prof_conv.py > gpurun_out/prof_conv_i.log 2>&1; tail -2 gpurun_out/prof_conv_i.log
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1i_step_eager.csv python bench.py --steps 1 --warmup 3 --no-graphs --no-overlap --no-cpu-baseline > gpurun_out/ncu_bench_i.log 2>&1; tail -2 gpurun_out/ncu_bench_i.log | cut -c1-200
