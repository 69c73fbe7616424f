set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; tail -c 300 gpurun_out/bench_r1j.err
cut -c1-260 gpurun_out/bench_r1j.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python scripts/graph_timeline.py 2>&1 | grep -v -i warn | head -12
