timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in 1 0 1 0; do SIVO_B200_PDL=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/tmp/err.log | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('PDL=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['segnet_ms_per_frame'])"; done
SIVO_B200_PDL=1 timeout 100 python tools/e2e_breakdown.py 2>&1 | grep "ms / frame"
SIVO_B200_PDL=0 timeout 100 python tools/e2e_breakdown.py 2>&1 | grep "ms / frame"
