set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; tail -c 300 gpurun_out/bench_r1k.err
cut -c1-260 gpurun_out/bench_r1k.json
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/conv_r1k -f python scripts/prof_conv.py > gpurun_out/prof_conv_k.log 2>&1; tail -2 gpurun_out/prof_conv_k.log
