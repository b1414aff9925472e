run() { echo "== $*"; env "$@" timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['conv_ms_per_frame'], d['roofline']['segnet_ms_per_frame'], d['e2e']['value'])"; }
run A=1
run SIVO_B200_TC_PAIR=0
run SIVO_B200_TC_NOEPI=1
run SIVO_B200_TC_NOEPI=1 SIVO_B200_TC_PAIR=0
run SIVO_B200_TC_NOEPI=3
run SIVO_B200_TC_NOEPI=2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
