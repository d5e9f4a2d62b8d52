timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r1C.json 2> gpurun_out/bench_r1C.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1C.json')); print('rqc30', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], {k:(v['launches_per_step'], round(v['ms_per_launch'],3)) for k,v in d['roofline']['kernels'].items()})"
python tools/e2e_profile.py --top 12 2>&1 | grep "e2e step"
