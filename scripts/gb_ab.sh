#!/bin/bash
# group-by (configs[3]) A/B runs: scripts/gb_ab.sh tag[:ENV=VAL,...] ...   -> gpurun_out/gb_<tag>.json + a one-line summary each
for spec in "$@"; do
  tag=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  env $envs timeout 300 python bench.py --query groupby --steps 5 --warmup 3 --no-cpu --no-e2e --extra-queries "" > gpurun_out/gb_$tag.json 2> gpurun_out/gb_$tag.err || tail -5 gpurun_out/gb_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/gb_%s.json" % tag).read().strip().splitlines()[-1])
    ks = d.get("kernels") or []
    print(tag, "ms/step %.2f" % d["ms_per_step"], "parity", (d.get("parity_check") or {}).get("ok"),
          " ".join("%s=%.2f" % (k["name"].replace("k_radix_", ""), k["total_ms"] / max(k["launches"], 1)) for k in ks))
except Exception as e:
    print(tag, "ERR", e)
PY
done
