for cfg in "256 5 0" "256 10 8" "256 3 0" "512 5 8" "512 3 16" "128 10 16" "256 20 4" "512 10 4"; do
  set -- $cfg
  echo "== T=$1 PPT=$2 CLUSTER=$3"
  DEMF_FPS_THREADS=$1 DEMF_FPS_PPT=$2 DEMF_FPS_CLUSTER=$3 python tools/kbench.py --only fps --reps 10 2>&1 | grep '"N": 20000\|"N": 2048' | cut -c1-160
done
