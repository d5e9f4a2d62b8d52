timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/sweep_bench.py --kind dm --qubits 10 --resolvers 256 --reps 1000 --out gpurun_out/sweep_r1x_dm10.json 2>&1 | tail -2
python tools/sweep_bench.py --kind sv --qubits 16 --resolvers 256 --reps 1000 --out gpurun_out/sweep_r1x_sv16.json 2>&1 | tail -2
python tools/sweep_bench.py --kind dm --qubits 12 --resolvers 32 --reps 1000 --seq-resolvers 8 --ref-resolvers 1 --out gpurun_out/sweep_r1x_dm12.json 2>&1 | tail -2
CIRQ_B200_TC_MODE=2 python bench.py --no-cpu-baseline > gpurun_out/bench_r1x_tc2.json 2> gpurun_out/bench_r1x_tc2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1x_tc2.json')); print('tc2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], {k:(v['launches_per_step'], round(v['ms_per_launch'],3)) for k,v in d['roofline']['kernels'].items()})"
