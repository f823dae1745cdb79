mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python scripts/parity_fuzz.py 200 8191 > gpurun_out/parity_fuzz_r2_seed8191_200cases.log 2>&1; tail -n 1 gpurun_out/parity_fuzz_r2_seed8191_200cases.log
