#!/bin/bash
# usage: tools/run_scaling.sh N TAG [extra bench.py args]   -- the driver's launch of bench.py at N GPUs, output under gpurun_out/
N=$1; TAG=$2; shift 2
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((30 + N)) bench.py --gpus $N "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
fi
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}.json"))
    print("${TAG}", "ms/step", round(d["ms_per_step"], 3), "Gvox/s", round(d["value"] / 1e9, 2), "verify", (d.get("verify") or {}).get("mismatching_ranks"),
          "classes", {k: round(v, 3) for k, v in d["roofline"]["per_class_ms_per_step"].items()}, "e2e ms", d["e2e"] and round(d["e2e"]["ms_per_step"], 1))
except Exception as e:
    print("${TAG} failed:", e); print(open("gpurun_out/${TAG}.err").read()[-1500:])
PY
