set -x
python scripts/microbench.py > gpurun_out/microbench_r2a.txt 2>&1; cat gpurun_out/microbench_r2a.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python scripts/prof_convr.py 55296 32 2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_buoyan_dilute" -s 1 -c 2 -o gpurun_out/prof_r2a -f python scripts/prof_convr.py 55296 32 2 > gpurun_out/ncu_r2a.log 2>&1; tail -3 gpurun_out/ncu_r2a.log
