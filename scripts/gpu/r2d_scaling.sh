# strong- and weak-scaling records on one 8-GPU box (N = 1, 2, 4, 8)
mkdir -p gpurun_out
run() {  # $1 = N, $2 = scaling, $3 = extra tag
  out=gpurun_out/bench_r2_f09_${2}_${1}gpu.json
  if [ "$1" = "1" ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --scaling $2 --no-cpu-baseline --no-extra-configs > $out 2> ${out%.json}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) \
      bench.py --gpus $1 --steps 20 --warmup 5 --scaling $2 > $out 2> ${out%.json}.err
  fi
  python -c "
import json,sys
d=json.load(open('$out'))
print('$2 N=$1', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,2),'M col/s  e2e', round(d['e2e']['value']/1e6,2), d.get('multi_gpu_parity',{}).get('equal'))" || tail -5 ${out%.json}.err
}
run 1 weak
for n in 2 4 8; do run $n strong; done
for n in 2 4 8; do run $n weak; done
