# final state: GPU suite, bench line, config 2 / 3, launch list, phases
set -x
T=r2_final2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_gputest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -c 300 gpurun_out/${T}_bench.err; cut -c1-200 gpurun_out/${T}_bench.json
for c in 2 3; do timeout 300 python bench.py --config $c > gpurun_out/${T}_config$c.json 2> gpurun_out/${T}_config$c.err; echo rc=$?; done
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python scripts/step_profile.py > gpurun_out/${T}_launches.log 2>&1; tail -2 gpurun_out/${T}_launches.log
SCDA_TIMESTAMPS=1 timeout 300 python scripts/phase_times.py > gpurun_out/${T}_phases.txt 2>&1; tail -16 gpurun_out/${T}_phases.txt
