PMB_LIB=build_exp/libpmb200_lanprof.so python tools/eig_prof.py 1000 6
for a in "1000 6" "2000 10" "5000 20" "700 3" "300 40"; do python tools/eig_bench.py $a 3; done 2>&1 | grep -v "rep [12]"
timeout 600 python -m pytest tests -x -q -m gpu -k "eig or c5 or implied or lanczos or enhanced or msm" 2>&1 | tail -3
