for cfg in "256 64 64 120 14 14" "64 48 40 90 16 16" "256 32 32 150 7 7" "512 16 16 30 14 14" "128 50 70 200 32 32" "4 9 5 40 3 2" "36 33 47 300 11 5" "260 20 20 40 7 7" "16 24 24 0 7 7" "8 33 33 50 1 1"; do
  echo "== $cfg"; timeout 60 python tools/dbg_bwd.py $cfg 2>&1 | tail -3; echo "rc=$?"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "crop_backward or autograd or pyramid_roi_align_matches" 2>&1 | tail -5
timeout 300 python tools/ab_bwd_impl.py 1000 4000 2>&1 | grep -v "^{\"peak" | tail -8
