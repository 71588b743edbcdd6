# Measurement helper (8-GPU box): the bench at N = 1, 2, 4, 8 (as the driver launches it) + config C5 data-parallel.
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-ref-gpu > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err; fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_scale_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.0f f/s (%.3f ms)  e2e %.0f f/s (%.3f ms)  parity %s  stages %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("parity",{}).get("ok"), {k: round(v,3) for k,v in d["roofline"]["stage_ms_per_step"].items()}))
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/r2_scale_n$n.err").read()[-1500:])
PY
done
for n in 1 8; do
  if [ $n = 1 ]; then python scripts/bench_train.py --batch 16 --steps 3 --cpu-batch 0 2>/dev/null | tail -1
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29650 scripts/bench_train.py --batch 16 --steps 3 --cpu-batch 0 2>/dev/null | tail -1; fi
done
