#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_colour.py -q -x --timeout 900 -k "carry" -s > gpurun_out/t_carry.log 2>&1
echo "carry tests rc=$?"; grep -E "as-shipped|passed|failed|Error|error" gpurun_out/t_carry.log | tail -12
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, "pathintegral-qmc_b200")
import piqmc.qmc as qmc
vec = np.load("tests/golden/ref_vectors.npz")
for inst, n in (("inst_0_32x32", 1024), ("santoro_80x80", 6400)):
    nbs = vec["nbs_" + inst]
    for R in (1024, 16384):
        sched = np.linspace(1.5, 1e-8, 100)
        qmc.QuantumAnnealReplicas(sched[:2], 1, 20, 0.01, n, None, nbs, 1, order="permutation", semantics="reference", nreplicas=R)
        out = qmc.QuantumAnnealReplicas(sched, 1, 20, 0.01, n, None, nbs, 1, order="permutation", semantics="reference", nreplicas=R)
        dt = out["seconds"]["sweeps"]
        print("%s carry mode R=%d P=20 100 steps: sweeps %.3f s -> %.3e attempts/s" % (inst, R, dt, R * 20.0 * n * 100 / dt))
PY
