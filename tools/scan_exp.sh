# needs the library built with: B200_NVCC_EXTRA=-DB200_TIMING_EXPERIMENTS python clip-retrieval_b200/build.py -f
for e in 0 1 2 0; do echo "EXP=$e"; B200_SCAN_EXP=$e timeout -k 10 120 python tools/knn_debug.py 30000000 2>&1 | grep "nq=8 \|nq=128\|nq=256\|nq=1000"; done
