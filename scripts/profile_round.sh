#!/bin/bash
# Evidence for profiles/: plain run, ncu launch list and one `ncu --set full` capture of the flood kernels on the
# bench workload (B200_PROFILING.md recipe).  Usage on the GPU box:  bash scripts/profile_round.sh r05
# Outputs go to gpurun_out/; summarise them into profiles/ with scripts/summarise_profile.py.
set -u
TAG=${1:-rXX}
CMD="python bench.py --loading 2 --steps 1 --warmup 1 --skip-cpu --skip-e2e --skip-extra"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-200
# every launch of one step with its device time (serialised, cold cache: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_fp16_P4096.csv \
    $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# the flood kernels: skip the first iterations, capture two CN + two VN launches
ncu --set full --clock-control none --import-source on -k regex:'cn_update_kernel|vn_update_kernel' -s 20 -c 4 \
    -f -o gpurun_out/${TAG}_prof_fp16_P4096 $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/${TAG}_*
