"""CPU oracle: test infrastructure only (see piqmc_oracle.c)."""
