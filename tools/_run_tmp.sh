timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/microbench.py --only diag --dense --reps 20 2>&1 | grep -E "diag|torch_copy"
python bench.py > gpurun_out/bench_r1A.json 2> gpurun_out/bench_r1A.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1A.json')); print('rqc30', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'], {k:(v['launches_per_step'], round(v['ms_per_launch'],3)) for k,v in d['roofline']['kernels'].items()})"
python bench.py --workload qft34 --no-cpu-baseline --steps 3 > gpurun_out/bench_r1A_qft34.json 2> gpurun_out/bench_r1A_qft34.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1A_qft34.json')); print('qft34', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['passes_per_step'], d['roofline']['frac'], d['roofline']['kernel'], {k:(v['launches_per_step'], round(v['ms_per_launch'],3)) for k,v in d['roofline']['kernels'].items()})"
( time python bench.py --impl reference --steps 1 --warmup 0 ) 2>&1 | tail -5 | cut -c1-900
