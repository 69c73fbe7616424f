# conv_first fix + whole GPU suite + kernel timeline of one replayed iteration
set -x
T=r2_r
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$?; tail -5 gpurun_out/${T}_gputest.log
timeout 300 python scripts/graph_timeline.py > gpurun_out/${T}_timeline.txt 2>&1; mv gpurun_out/timeline.tsv gpurun_out/${T}_timeline.tsv; head -4 gpurun_out/${T}_timeline.txt
