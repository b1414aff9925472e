timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/tmp/err.log | tail -1 > gpurun_out/bench_r1f.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1f.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
r=d['roofline']; print(r['frac'], r['kernel_ms'], r['all_conv_launches'], r['segnet_ms_per_frame']); print(r['launch_ms'])
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --model standard 2>/tmp/err.log | tail -1 > gpurun_out/bench_r1f_std.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1f_std.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
r=d['roofline']; print(r['frac'], r['kernel_ms'], r['all_conv_launches'], r['segnet_ms_per_frame']); print(r['launch_ms'])
PY
