timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in 1 4 8 4 1; do echo "bands=$b"; SIVO_B200_READBACK_BANDS=$b timeout 100 python tools/e2e_breakdown.py 2>&1 | grep "ms / frame" | grep -v "left on\|alone   " ; done
for b in 1 4; do SIVO_B200_READBACK_BANDS=$b timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('bands=$b', d['value'], d['e2e']['value'])"; done
