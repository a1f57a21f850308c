for d in 1 2 3 4 5 6 0; do
  echo "== dbg $d"; NCA_T2_DBG=$d timeout 120 python tools/dbg_tc2.py ${1:-1} 2>&1 | grep -E "variant|rel err|Error|error" | head -3
done
