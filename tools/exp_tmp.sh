timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
run() { echo "== $*"; env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/tmp/err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_frame'], d['roofline']['segnet_ms_per_frame'], d['e2e']['value'], d['host_ms_per_step'])"; }
run A=1
run A=1
run SIVO_B200_TC_NOEPI=3
run SIVO_B200_TC_NOEPI=1
