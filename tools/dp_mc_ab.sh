# A/B of Row G: NCCL two-phase allreduce + replicated Adam  vs  the fused multicast kernel (csrc/dp_update.cu).
# usage: dp_mc_ab.sh <ranks> "name:ENV=V,ENV=V[:bench flags]" ...
N=$1; shift
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; envs=${rest%%:*}; flags=""; [ "$rest" != "$envs" ] && flags=${rest#*:}
  if [ "$N" = 1 ]; then launcher="python"; else launcher="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
  env $(echo $envs | tr ',' ' ') timeout 300 $launcher bench.py --gpus $N --steps 100 --warmup 10 --no-extras --no-cpu-baseline --no-e2e $flags > gpurun_out/dpmc${N}_$name.json 2> gpurun_out/dpmc${N}_$name.err
  echo "== N=$N $name rc=$? $(grep -o '"ms_per_step": [0-9.]*, "ms_per_step_ranks": \[[^]]*\]' gpurun_out/dpmc${N}_$name.json | head -1)"
done
