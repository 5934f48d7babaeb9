for e in 0 1 2 0; do echo "EXP=$e"; B200_SCAN_EXP=$e timeout -k 10 120 python tools/knn_debug.py 30000000 2>&1 | grep "nq=8 \|nq=128\|nq=256\|nq=1000"; done
