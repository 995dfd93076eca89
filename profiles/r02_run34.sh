#!/bin/bash
# final evidence pass of round 2
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02P_pytest_gpu.log
grep -E "Error|assert|passed|failed|FAILED" gpurun_out/r02P_pytest_gpu.log | head -30
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02P_bench.json 2>> gpurun_out/r02P.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02P_bench_steps20.json 2>> gpurun_out/r02P.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02P_bench_reference.json 2>> gpurun_out/r02P.err
timeout 600 python bench_configs.py --configs c1,c3,c4,c5rbm --reps 3 > gpurun_out/r02P_configs.jsonl 2>> gpurun_out/r02P.err
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 800 --csv --log-file gpurun_out/r02P_launches_bench_steps20.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02P_bench_under_ncu.json 2>> gpurun_out/r02P_ncu.err
$NCU --set full --import-source on -k "regex:fc_warp_mc_kernel" --launch-skip 2 -c 1 -f -o /tmp/ncu/r02P_fc_warp python bench_configs.py --configs c1 --reps 1 > /dev/null 2>> gpurun_out/r02P_ncu.err
python profiles/summarize_ncu.py /tmp/ncu/r02P_fc_warp.ncu-rep > gpurun_out/r02P_fc_warp_ncu_full.txt 2>> gpurun_out/r02P_ncu.err
python profiles/source_hotspots.py /tmp/ncu/r02P_fc_warp.ncu-rep fc_warp_mc_kernel 25 > gpurun_out/r02P_fc_warp_hotspots.txt 2>> gpurun_out/r02P_ncu.err
tail -3 gpurun_out/r02P_ncu.err
python - <<'PY'
import json
for f in ('gpurun_out/r02P_bench.json','gpurun_out/r02P_bench_steps20.json'):
  for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print(f, {k:d.get(k) for k in ('value','ms_per_step','step_ms','epoch_end_ms','steps')}, 'roof', round(d['roofline']['frac'],3))
        print('  e2e', e['ms_per_step'], 'host_pack', e['host_pack'], 'other', e['other_upload_form']['ms_per_step'])
        for k,v in d['configs'].items(): print('  ',k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
tail -3 gpurun_out/r02P.err
echo done
