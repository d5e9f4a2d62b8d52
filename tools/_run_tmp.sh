timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1t.json 2> gpurun_out/bench_r1t.err; tail -c 2500 gpurun_out/bench_r1t.json
python bench.py --workload qft30 --no-cpu-baseline > gpurun_out/bench_r1t_qft30.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_r1t_qft30.json')); print('qft30', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1t.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_r1t.log 2>&1; tail -2 gpurun_out/ncu_launch_r1t.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:sv_apply_tc_staged -c 3 -o gpurun_out/prof_tcs_r1t python tools/microbench.py --only k5_low,k5_bit1_mid --dense --reps 1 > gpurun_out/ncu_tcs_r1t.log 2>&1; tail -2 gpurun_out/ncu_tcs_r1t.log | cut -c1-200
