timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/traj_bench.py --qubits 16 --depth 8 --reps 4096 --batch 4096 --out gpurun_out/traj_r1w_16q.json 2>&1 | tail -2
python tools/traj_bench.py --qubits 10 --depth 8 --reps 65536 --batch 65536 --loop-reps 128 --ref-reps 64 --out gpurun_out/traj_r1w_10q.json 2>&1 | tail -2
python tools/traj_bench.py --qubits 20 --depth 8 --reps 1024 --batch 1024 --loop-reps 32 --ref-reps 2 --out gpurun_out/traj_r1w_20q.json 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err; tail -c 1500 gpurun_out/bench_r1w.json
