for cfg in "" "B3D_NOPATCH=1" "B3D_TH=16 B3D_TW=8" "B3D_TH=16 B3D_TW=16" "B3D_TH=16 B3D_TW=32" "B3D_TH=32 B3D_TW=64" "B3D_NOPATCH=1 B3D_TH=1" "B3D_NOPATCH=1 B3D_TH=2"; do
  echo "== $cfg"; env $cfg B3D_VERBOSE=1 timeout 120 python scripts/gpu_check_conv.py pw 2>&1 | grep -E "^conv|igemm" | sort | uniq | cut -c1-185
done
