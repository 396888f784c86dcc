for v in libpshadow_prev.so libpshadow.so libpshadow_prev.so libpshadow.so; do
PSH_LIB=$v timeout 300 python tests/prof_scan_sizes.py 32768 2>&1 | sed "s/^/$v /"
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "seedless or fft_flavours or golden" 2>&1 | tail -2
timeout 100 python tests/timeline.py 2>&1 | sed -n 3,5p
