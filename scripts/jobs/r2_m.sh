# round-2 evidence job (1 GPU): GPU suite, bench line, timeline, operator bench, launch list of one
# replayed iteration, ncu --set full of the RoI / NMS operators
set -x
T=r2_m
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; tail -3 gpurun_out/${T}_gputest.log
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
python scripts/graph_timeline.py > gpurun_out/${T}_timeline.txt 2>&1; mv gpurun_out/timeline.tsv gpurun_out/${T}_timeline.tsv; head -4 gpurun_out/${T}_timeline.txt
python scripts/opbench.py > gpurun_out/${T}_opbench.jsonl 2> gpurun_out/${T}_opbench.err; tail -3 gpurun_out/${T}_opbench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python scripts/step_profile.py > gpurun_out/${T}_launches.log 2>&1; tail -2 gpurun_out/${T}_launches.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'roi_|nms_' -c 14 -f -o gpurun_out/${T}_ops python scripts/prof_ops.py all > gpurun_out/${T}_ops.log 2>&1; tail -3 gpurun_out/${T}_ops.log
