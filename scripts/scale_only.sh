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
