for c in 12 13 14; do
  echo "== WINDOW $c"
  PCD_MSM_WINDOW=$c PROBE=main,help python tools/probe_pcd.py 2>&1 | grep -E "^==|acc_g2|reduce"
done
