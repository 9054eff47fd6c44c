for a in "1000 6" "2000 10" "5000 20"; do python tools/eig_bench.py $a 4; done 2>&1 | grep -v "rep [12]"
timeout 600 python -m pytest tests -x -q -m gpu -k "eig or c5 or implied or lanczos or enhanced" 2>&1 | tail -3
