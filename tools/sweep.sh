run() { echo "$@"; env "$@" timeout 120 python bench.py --no-cpu --steps 50 $EXTRA 2>&1 | grep -o '"value": [0-9.e+]*' | head -1; }
timeout 600 python -m pytest tests/test_gpu_colour.py -m gpu -x -q 2>&1 | tail -2
PIQMC_FORCE_GENERIC_FN=1 timeout 300 python -m pytest tests/test_gpu_colour.py -m gpu -x -q -k "bit_exact or world or periodic" 2>&1 | tail -1
EXTRA=""
run PIQMC_MINB=9
run PIQMC_MINB=8
EXTRA="--replicas 512"
run PIQMC_MINB=9
