# needs the library built with: B200_NVCC_EXTRA=-DB200_TIMING_EXPERIMENTS python clip-retrieval_b200/build.py -f
# A/B of the pair GEMM under sustained clocks: baseline, epilogue without global stores, without the TMEM read-out,
# with half of B loaded.
for v in "" "B200_GEMM_NOSTORE=1" "B200_GEMM_NOLDTM=1" "B200_GEMM_HALFB=1"; do
  echo "== ${v:-baseline}"
  env $v timeout -k 10 200 python tools/gemm_sustained.py 2>&1 | grep "TFLOP"
done
