python tools/e2e_profile.py --top 45 2>&1 | cut -c1-170 | tail -70
python tools/sweep_bench.py --kind sv --qubits 16 --resolvers 256 --expect --out gpurun_out/sweep_r1B_sv16_expect.json 2>&1 | tail -1 | cut -c1-1200
python tools/sweep_bench.py --kind dm --qubits 10 --resolvers 256 --expect --ref-resolvers 1 --out gpurun_out/sweep_r1B_dm10_expect.json 2>&1 | tail -1 | cut -c1-1200
