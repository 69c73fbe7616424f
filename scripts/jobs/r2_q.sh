# state check at 9c80ed1 (1 GPU): GPU suite, bench line, smoke, per-layer convolution / weight-gradient bench,
# launch list of one replayed iteration
set -x
T=r2_q
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$?; tail -5 gpurun_out/${T}_gputest.log
timeout 420 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -c 400 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/convbench.py > gpurun_out/${T}_convbench.jsonl 2> gpurun_out/${T}_convbench.err; echo rc=$?; tail -3 gpurun_out/${T}_convbench.err
timeout 300 python scripts/halobench.py > gpurun_out/${T}_halobench.jsonl 2> gpurun_out/${T}_halobench.err; echo rc=$?; tail -3 gpurun_out/${T}_halobench.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python scripts/step_profile.py > gpurun_out/${T}_launches.log 2>&1; tail -2 gpurun_out/${T}_launches.log
SCDA_TIMESTAMPS=1 timeout 200 python scripts/phase_times.py > gpurun_out/${T}_phases.txt 2>&1; tail -30 gpurun_out/${T}_phases.txt
