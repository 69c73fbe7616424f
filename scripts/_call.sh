set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1l.json 2> gpurun_out/bench_r1l.err; tail -c 300 gpurun_out/bench_r1l.err
cut -c1-260 gpurun_out/bench_r1l.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 60 2>/dev/null | cut -c1-400
