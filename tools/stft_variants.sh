#!/bin/bash
set -u
echo "== pytest"; timeout 600 python -m pytest tests/test_gpu_stft.py -q 2>&1 | tail -4
for cfg in "AA_STFT_V2_MEL=1 AA_STFT_DIAG=0" "AA_STFT_V2_MEL=1 AA_STFT_DIAG=1"; do
  echo "== time $cfg"; env $cfg MODES=mel timeout 300 python tools/time_stft.py 2>&1 | tail -22
done
