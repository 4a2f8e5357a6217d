for mw in 16 10 8 5 4 3; do echo "MAXW=$mw"; DSSMD_MMA16_MAXW=$mw python scripts/dtype_bench.py 2>&1 | grep "bfloat16"; done
