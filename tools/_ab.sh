for c in 0 4 5 6 7 8; do
  echo "== WINDOW $c"
  PCD_MSM_WINDOW=$c PROBE=tiny python tools/probe_pcd.py 2>&1 | grep -E "^== tinypre"
done
