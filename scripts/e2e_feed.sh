#!/bin/bash
# Host-feed experiment: the e2e leg of bench.py with different numbers of host threads for the staging gather
# (a rank has hardware_concurrency / ranks threads; 4 on the 8-GPU, 32-core box).  Usage: bash scripts/e2e_feed.sh TAG 16 4 2
TAG=$1; shift
for t in "$@"; do
  LDPC_B200_HOST_THREADS=$t python bench.py --steps 1 --warmup 1 --skip-cpu --skip-extra --small-pool 0 > gpurun_out/${TAG}_threads$t.log 2>&1
  tail -1 gpurun_out/${TAG}_threads$t.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('threads $t: value %.1f e2e %.1f Mb/s  step %.3f vs %.3f s  host_stage %.2f s (first batch %.2f s)  device %.2f s' % (d['value'], e['value'], e['ms_per_step']/1e3, d['ms_per_step']/1e3, e['host_stage_s_per_step'], e['first_batch_stage_s'], e['device_s_per_step']))"
done
