#!/bin/bash
# A/B of whole bench steps: one `bench.py --steps 1 --warmup 1` per (library, environment) pair, same box, back to back.
# Usage: bash scripts/ab_bench.sh TAG "lib_a_base.so:" "lib_b_mb3.so:LDPC_B200_FUSED_REFILL=1" ...
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  lib=${spec%%:*}; envs=${spec#*:}
  name=$(echo "${lib%.so}_${envs}" | tr ' =' '__')
  echo "### $lib $envs"
  env LDPC_B200_LIB=ab/$lib $envs python bench.py --steps 1 --warmup 1 --skip-e2e --skip-cpu --skip-extra > gpurun_out/${TAG}_${name}.log 2>&1
  tail -1 gpurun_out/${TAG}_${name}.log | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read())
    r=d['roofline']; s=r['service_s_per_step']
    print('value %.1f Mb/s step %.3f s | CN %.0f VN %.0f pair %.4f | service check %.3f retire %.3f refill %.3f cn_refill %.3f (%d fused passes) | avg_iter %.2f clocks %s' % (d['value'], d['ms_per_step']/1e3, r['cn']['gbs'], r['vn']['gbs'], r['frac'], s['check_s'], s['retire_s'], s['refill_s'], s['cn_refill_s'], r['cn_refill_launches'], d['avg_iter'], d['clocks']['sm_mhz']))
except Exception as e:
    print('failed', e)
"
done
