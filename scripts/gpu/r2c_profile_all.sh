set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/prof_all.py 55296 2 > gpurun_out/plain_all.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"^k_" -o gpurun_out/prof_r2b_all -f python scripts/prof_all.py 55296 1 > gpurun_out/ncu_r2b.log 2>&1
tail -3 gpurun_out/plain_all.log gpurun_out/ncu_r2b.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; tail -c 300 gpurun_out/bench_r2g.err
