set -x
for i in 1 2; do
python tools/km_bench.py c4 2>&1 | grep "^c4 n=" 
PMB_LIB=build_exp/libpmb200_kmhead.so python tools/km_bench.py c4 2>&1 | grep "^c4 n="
done
