# Scratch script for one-off GPU experiments (run through gpurun from the repo root).  The micro-benchmarks
# of the round, each printing one line per configuration:
python tools/prj_bench.py            # K5 projection, 10 M x 256 -> 10
python tools/tica_bench.py 256       # K4 TICA solve (PMB_TICA_CHOL=0 / PMB_TICA_CLUSTER=0 for the A/B forms)
python tools/eig_bench.py 1000 6 4   # K9 Lanczos, top-6 of a 1000-state reversible chain
python tools/km_bench.py c4 2>&1 | tail -5   # K6 k-means assignment: cold / subsampled hints / warm / with accumulation
