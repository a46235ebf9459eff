"""Instance access for the example scripts: the reference's instance files are stored (as data) in
tests/golden/instances.npz together with the Spin-Glass-Server ground states."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pathintegral-qmc_b200"))
_DATA = np.load(os.path.join(ROOT, "tests", "golden", "instances.npz"))


def load(name, nspins):
    import piqmc.tools as tools
    return tools.IsingFromTriples(_DATA["inst_" + name], nspins)


def ground_state(name):
    return _DATA["gs_" + name].astype(np.float64), float(_DATA["gs_energy_cie_" + name])
