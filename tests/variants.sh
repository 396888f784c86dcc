#!/bin/bash
# timing of the fft scan variants through bench.py (no cpu leg): "LIB QREG REFRESH"
run() {
  echo "== LIB=$1 QREG=$2 REFRESH=$3"
  PSH_LIB=$1 PSH_FFT_QREG=$2 PSH_FFT_REFRESH=$3 python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('ms_per_step',round(d['ms_per_step'],4),'scan_ms',round(r['kernel_ms_per_step'],4),'frac',round(r['frac'],3),'select_ms',round(r['select_ms_per_step'],4),'e2e_ms',round(d['e2e']['ms_per_step'],4),'launches',d['gpu_launches'],'host_ms',round(d['host_enqueue_ms_per_step'],4))"
}
while read -r line; do [ -n "$line" ] && run $line; done <<< "$VARIANTS"
