run() { echo "$@"; env "$@" timeout 120 python bench.py --no-cpu --steps 50 $EXTRA 2>&1 | grep -o '"value": [0-9.e+]*' | head -1; }
EXTRA=""
run PIQMC_BLOCKS=592
run PIQMC_BLOCKS=888
run PIQMC_BLOCKS=1036
EXTRA="--replicas 512"
run PIQMC_BLOCKS=444
run PIQMC_BLOCKS=592
run PIQMC_BLOCKS=888
