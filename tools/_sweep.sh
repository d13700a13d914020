for L in 3 4 5; do echo "== LOGL $L"; PCDGPU_REDUCE_LOGL=$L PROBE=main,help PROBE_OUT=probe_l$L.json python tools/probe_pcd.py 2>&1 | grep -E "^==|reduce|assemble"; done
