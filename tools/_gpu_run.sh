N=${NGPU:-8}
for wl in jbu loftup train; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl > gpurun_out/n${N}_$wl.json 2> gpurun_out/n${N}_$wl.err
  echo "== $wl rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/n${N}_$wl.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), d['scaling'], d['clocks'])"
done
