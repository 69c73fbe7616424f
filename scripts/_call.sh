set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; tail -c 300 gpurun_out/bench_r1m.err
cut -c1-260 gpurun_out/bench_r1m.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
