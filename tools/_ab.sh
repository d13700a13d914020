for ctas in 2 3; do
echo "=== ACC_CTAS $ctas"
PCDGPU_ACC_CTAS=$ctas PROBE=main,help python tools/probe_pcd.py 2>&1 | grep -E "^=="  | cut -c1-110
PCDGPU_ACC_CTAS=$ctas PROBE=main,help python tools/probe_pcd.py 2>&1 | grep -E "^=="  | cut -c1-110
done
