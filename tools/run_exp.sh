for s in 2 3 4 6; do PMB_PRJ_STAGES=$s PMB_PRJ_VARIANT=5 python tools/prj_bench.py; done
PMB_PRJ_VARIANT=5 python tools/prj_bench.py 10000000 128 4
PMB_PRJ_VARIANT=5 python tools/prj_bench.py 10000000 256 16
PMB_PRJ_VARIANT=5 python tools/prj_bench.py 9999999 84 3
PMB_PRJ_VARIANT=5 timeout 600 python -m pytest tests -x -q -m gpu -k "project or tica or preprocess" 2>&1 | tail -3
