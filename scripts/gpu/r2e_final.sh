set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python scripts/parity_fuzz.py 60 2027 > gpurun_out/parity_fuzz_r2_seed2027_60cases.log 2>&1; tail -3 gpurun_out/parity_fuzz_r2_seed2027_60cases.log
python scripts/prof_all.py 55296 2 > gpurun_out/plain_all.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_buoyan_dilute" -o gpurun_out/prof_r2c_cape -f python scripts/prof_all.py 55296 1 > gpurun_out/ncu_r2c.log 2>&1
tail -2 gpurun_out/ncu_r2c.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_f09_final.json 2> gpurun_out/bench_r2_f09_final.err; tail -c 300 gpurun_out/bench_r2_f09_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null
python scripts/microbench.py > gpurun_out/microbench_r2_final.txt 2>&1
