run() { echo "$@"; env "$@" timeout 120 python bench.py --no-cpu --steps 50 $EXTRA 2>&1 | grep -o '"value": [0-9.e+]*' | head -1; }
EXTRA=""
run PIQMC_ROWS_PER_BLOCK=1024
run PIQMC_ROWS_PER_BLOCK=768
run PIQMC_ROWS_PER_BLOCK=384
EXTRA="--replicas 2048"
run PIQMC_ROWS_PER_BLOCK=512
run PIQMC_ROWS_PER_BLOCK=384
run PIQMC_ROWS_PER_BLOCK=256
EXTRA="--replicas 1024"
run PIQMC_ROWS_PER_BLOCK=256
run PIQMC_ROWS_PER_BLOCK=128
EXTRA="--replicas 512"
run PIQMC_ROWS_PER_BLOCK=128
run PIQMC_ROWS_PER_BLOCK=256
