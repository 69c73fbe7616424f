# round-2 final evidence job (1 GPU): GPU suite, bench line (+ reference arm, short budget), configs 2/3/5, operator /
# convolution benches, launch list of one replayed iteration, ncu --set full of the convolution / weight-gradient /
# RoI / NMS kernels, compute-sanitizer memcheck over the operator parity tests, smoke
set -x
T=r2_w
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_gputest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -c 300 gpurun_out/${T}_bench.err; cut -c1-200 gpurun_out/${T}_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 --cpu-budget 60 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo rc=$?; cut -c1-400 gpurun_out/${T}_bench_ref.json
for c in 2 3 5; do timeout 300 python bench.py --config $c > gpurun_out/${T}_config$c.json 2> gpurun_out/${T}_config$c.err; echo rc=$?; cut -c1-300 gpurun_out/${T}_config$c.json; done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/opbench.py > gpurun_out/${T}_opbench.jsonl 2> gpurun_out/${T}_opbench.err; echo rc=$?
timeout 300 python scripts/convbench.py > gpurun_out/${T}_convbench.jsonl 2> gpurun_out/${T}_convbench.err; echo rc=$?
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python scripts/step_profile.py > gpurun_out/${T}_launches.log 2>&1; tail -2 gpurun_out/${T}_launches.log
for p in fwd dgrad wgrad; do timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/${T}_conv_$p python scripts/prof_conv.py --pass=$p > gpurun_out/${T}_conv_$p.log 2>&1; tail -1 gpurun_out/${T}_conv_$p.log; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'roi_|nms_' -c 14 -f -o gpurun_out/${T}_ops python scripts/prof_ops.py all > gpurun_out/${T}_ops.log 2>&1; tail -1 gpurun_out/${T}_ops.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ops_gpu.py tests/test_roi_pool_nhwc_gpu.py -x -q > gpurun_out/${T}_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/${T}_memcheck.log
ls -la gpurun_out/*.ncu-rep
