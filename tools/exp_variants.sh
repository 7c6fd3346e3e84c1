#!/bin/bash
# Developer experiment driver (runs on the GPU box): rebuild with compile-time variants and time each.
# usage: tools/exp_variants.sh "ENV1=.. EXTRA='-D..'" ...   each argument is one variant: "name|env assignments|nvcc extra"
for v in "$@"; do
  name="${v%%|*}"; rest="${v#*|}"; envs="${rest%%|*}"; extra="${rest#*|}"
  echo "=== $name  env[$envs] extra[$extra]"
  env $envs P252_NVCC_EXTRA="$extra" python -m poseidon252_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  grep -E "Used" poseidon252_b200/lib/build.log | sort | uniq -c | head -3
  python tools/quick_bench.py 2>&1 | head -2
done
