#!/bin/bash
# timing of fft scan knobs through bench.py (no cpu leg): each line of $VARIANTS is a list of VAR=value
run() {
  echo "== $*"
  env "$@" python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('ms_per_step',round(d['ms_per_step'],4),'scan_ms',round(r['kernel_ms_per_step'],4),'frac',round(r['frac'],3),'select_ms',round(r['select_ms_per_step'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),'parity',d['parity_checked'])"
}
while read -r line; do [ -n "$line" ] && run $line; done <<< "$VARIANTS"
