#!/bin/bash
# usage: tools/gpu_sweep.sh "inflight:opt=val,opt=val" ...   -> one bench line summary per config
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  n=${cfg%%:*}; o=${cfg#*:}
  opts=""; IFS=',' read -ra kv <<< "$o"; for x in "${kv[@]}"; do [ -n "$x" ] && opts="$opts --opt $x"; done
  timeout 200 python bench.py --inflight $n --steps ${STEPS:-6} --no-cpu-baseline $opts > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  python - "$i" "$cfg" <<'PY'
import json,sys
i,cfg=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f'gpurun_out/sweep_{i}.json').read().strip().splitlines()[-1])
    p=d['phase_ms']
    print(f"{cfg:40s} value {d['value']:7.2f} e2e {d['e2e']['value']:7.2f} ms {d['ms_per_step']:6.1f} enc {p['encode_ms']:.1f} pt {p['pt_loop_ms']:.1f} pr {p['poly_rec_loops_ms']:.1f}")
except Exception as e:
    print(cfg, 'ERR', e); print(open(f'gpurun_out/sweep_{i}.err').read()[-1200:])
PY
done
