"""piqmc on B200: drop-in for the reference package `piqmc` (hadsed/pathintegral-qmc).

    import piqmc.sa as sa, piqmc.qmc as qmc, piqmc.tools as tools

`tools` is pure host code and imports without the CUDA library; `sa`, `qmc` and `device`
bind libpiqmc_b200.so and fail loudly (ImportError) if it has not been built.
"""
__version__ = "1.0.0"
