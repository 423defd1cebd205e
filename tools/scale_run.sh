# usage (on the GPU box): bash tools/scale_run.sh "1 2 4 8"
mkdir -p gpurun_out
NS=${1:-"1 2"}
timeout 300 python -m pytest tests/test_multirank_gloo.py -m gpu -q > gpurun_out/mg_test.log 2>&1; tail -3 gpurun_out/mg_test.log
for n in $NS; do
  if [ $n -eq 1 ]; then timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no_cpu_baseline --no_kernel_pass > gpurun_out/scale_$n.log 2>gpurun_out/scale_$n.err;
  else timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $n --steps 10 --warmup 3 --no_cpu_baseline --no_kernel_pass > gpurun_out/scale_$n.log 2>gpurun_out/scale_$n.err; fi
  echo "n=$n rc=$?"; tail -2 gpurun_out/scale_$n.err | cut -c1-300
  python -c "
import json
for l in open('gpurun_out/scale_$n.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done
