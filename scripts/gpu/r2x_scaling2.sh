# end-of-round check of the multi-GPU path on 2 GPUs (final build): smoke, strong and weak
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
run() {
  out=gpurun_out/bench_r2x_f09_${2}_${1}gpu.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) \
      bench.py --gpus $1 --steps 20 --warmup 5 --scaling $2 --no-cpu-baseline --no-extra-configs > $out 2> ${out%.json}.err
  python -c "
import json,sys
d=json.load(open('$out'))
print('$2 N=$1', round(d['ms_per_step'],3),'ms', round(d['value']/1e6,2),'M col/s  e2e', round(d['e2e']['value']/1e6,2), d.get('multi_gpu_parity',{}).get('equal'))" || tail -5 ${out%.json}.err
}
run 2 strong
run 2 weak
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
