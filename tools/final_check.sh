# round-end validation on one B200: GPU tests, smoke(), the bench line of record
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['all_conv_launches']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['clocks'])"
