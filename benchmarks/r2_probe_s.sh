#!/bin/bash
O=gpurun_out/r2s; mkdir -p $O
timeout 300 python benchmarks/linear_shapes.py 2> $O/linear_shapes.err | tail -1 > $O/linear_shapes.json
python - <<'PY'
import json
for r in json.load(open("gpurun_out/r2s/linear_shapes.json")):
    print(r["layer"], r["K"], r["N"], {k: r[k] for k in ("split", "split_pair", "fp16")})
PY
tail -3 $O/linear_shapes.err
