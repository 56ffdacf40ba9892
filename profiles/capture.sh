#!/bin/bash
# Produces the files behind profiles/<tag>_*: the bench line, the ncu launch list of one step (profiles/prof_step.py; one
# pass, no replay) and the full-set report of the Stage-1 launches (profiles/prof_tsdf.py, ~1 min).  Run on the GPU box from
# the repo root:
#     gpurun --timeout 330 -- 'bash profiles/capture.sh r2l'
# The full-set report of the whole step (every kernel replayed ~45 times with > 10 GB of device memory saved and restored
# around each pass) takes about a quarter of an hour and a report larger than gpurun's 64 MiB return limit with
# --import-source on; run it on its own, without sources:
#     gpurun --timeout 1500 -- 'ncu --profile-from-start off --set full --clock-control none -f -o gpurun_out/<tag> python profiles/prof_step.py'
# then, here:  ncu -i gpurun_out/<tag>.ncu-rep --page raw --csv > /tmp/raw.csv && python profiles/summarise_ncu.py /tmp/raw.csv profiles/<tag>_full_summary.csv
tag=${1:-run}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<EOF
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print(d["ms_per_step"], d["value"], d["e2e"])
EOF
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python profiles/prof_step.py > gpurun_out/ncu_${tag}_2.log 2>&1; tail -1 gpurun_out/ncu_${tag}_2.log
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/${tag}_tsdf python profiles/prof_tsdf.py > gpurun_out/ncu_${tag}_3.log 2>&1; tail -1 gpurun_out/ncu_${tag}_3.log
ls -la gpurun_out/
