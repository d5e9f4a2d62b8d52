timeout 900 python -m pytest tests -m gpu -x -q -k "batched or trajector" 2>&1 | tail -3
python tools/traj_bench.py --qubits 16 --depth 8 --reps 4096 --batch 4096 --out gpurun_out/traj_r1E_16q.json 2>&1 | tail -1 | cut -c1-700
python tools/traj_bench.py --qubits 10 --depth 8 --reps 65536 --batch 65536 --loop-reps 128 --ref-reps 64 --out gpurun_out/traj_r1E_10q.json 2>&1 | tail -1 | cut -c1-500
python tools/traj_bench.py --qubits 20 --depth 8 --reps 1024 --batch 1024 --loop-reps 32 --ref-reps 2 --out gpurun_out/traj_r1E_20q.json 2>&1 | tail -1 | cut -c1-500
