timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload qft34 --no-cpu-baseline --steps 3 > gpurun_out/bench_r1G_qft34.json 2> gpurun_out/bench_r1G_qft34.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1G_qft34.json')); print('qft34', d['value'], d['ms_per_step'], 'e2e', d['e2e'], d['config']['passes_per_step'], d['roofline']['frac'], d['roofline']['kernel'], {k:(v['launches_per_step'], round(v['ms_per_launch'],3)) for k,v in d['roofline']['kernels'].items()})"
tail -3 gpurun_out/bench_r1G_qft34.err
