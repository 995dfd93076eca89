python -m pytest tests/test_gpu_rbm.py -x -q -k "host_fed or graphed" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01G_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:walker_kernel --launch-skip 6 -c 1 -f -o gpurun_out/r01G_fused python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01G_ncu.log 2>&1
tail -1 gpurun_out/r01G_ncu.log | cut -c1-100
python bench.py > gpurun_out/r01G_bench.json 2> gpurun_out/r01G_bench.err
tail -c 300 gpurun_out/r01G_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r01G_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['packed_host_input']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'])"
python bench_configs.py --configs c5rbm --reps 3 --walker-sweep 16384,32768,65536,131072 > gpurun_out/r01G_c5_sweep.jsonl 2>/dev/null
python bench_configs.py --configs c1 --reps 5 >> gpurun_out/r01G_c5_sweep.jsonl 2>/dev/null
wc -l gpurun_out/r01G_c5_sweep.jsonl
