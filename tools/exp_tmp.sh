timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/tmp/err.log | tail -1 > gpurun_out/bench_r1g.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1g.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['orb_last_call'])
r=d['roofline']; print(r['frac'], r['kernel_ms'], r['all_conv_launches']['frac'], r['segnet_ms_per_frame']); print(r['launch_ms'])
PY
done
SIVO_B200_ORB_TREE_THREADS=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/tmp/err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1 thread', d['value'], d['e2e']['value'], d['orb_last_call'])"
