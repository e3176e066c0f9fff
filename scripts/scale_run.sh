#!/bin/bash
# The driver's scaling sequence on ONE multi-GPU box: bench.py at N = 1, 2, 4, 8 back to back (one rank per GPU under torchrun),
# each JSON line kept as gpurun_out/r02_scale_n<N>.json; then the library's own sharding check (ngpus = G against ngpus = 1).
#   gpurun --gpus 8 --timeout 1500 -- 'bash scripts/scale_run.sh'
mkdir -p gpurun_out
STEPS=${STEPS:-20}
python bench.py --gpus 1 --steps $STEPS --warmup 3 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
    bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
done
python scripts/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_8gpu.txt 2>&1
nvidia-smi topo -m > gpurun_out/r02_8gpu_topo.txt 2>&1
for N in 1 2 4 8; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_scale_n$N.json").read().strip().splitlines()[-1])
    s = d.get("secondary", {})
    print($N, "value %.3e" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"], "C2 %.3e" % s.get("value", 0), "C2 ms %.2f" % s.get("ms_per_step", 0),
          "large %.3e" % d.get("large", {}).get("value", 0), "weak %.3e" % d.get("weak", {}).get("value", 0), "1proc %.3e" % d.get("e2e_one_process", {}).get("value", 0))
except Exception as e:
    print($N, "failed", e)
PY
done
tail -3 gpurun_out/r02_multi_gpu_check_8gpu.txt
