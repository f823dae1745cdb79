mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong --no-cpu-baseline --no-extra-configs > gpurun_out/bench_r3j_strong2.json 2> gpurun_out/bench_r3j_strong2.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r3j_strong2.json'));print(d['ms_per_step'], d['multi_gpu_parity'])" || tail -5 gpurun_out/bench_r3j_strong2.err
