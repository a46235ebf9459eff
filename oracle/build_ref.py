#!/usr/bin/env python
"""Build the UNMODIFIED-arithmetic reference (hadsed/pathintegral-qmc) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path may import oracle/.

The reference is three Cython modules (piqmc/{sa,qmc,tools}.pyx).  As shipped they do not
build with Cython 3 / NumPy 2 / Python 3, for three reasons that have nothing to do with
arithmetic:

  * ``np.int_t`` no longer exists in NumPy 2's .pxd          (qmc.pyx:90,194; sa.pyx:91,158,336)
  * a Python-2 ``print svec_g.get()`` statement               (sa.pyx:613, inside Anneal_cuda)
  * ``J.iterkeys()`` is gone on Python 3                      (tools.pyx:84)

This recipe copies the three .pyx files from where they lie under /root/reference into a
scratch directory under /tmp, applies those three one-token shims with ``re.sub``,
cythonizes with language_level=2 and compiles with /usr/bin/gcc -O2 -fopenmp.  Only the
resulting extension modules are written into oracle/_ref/piqmc_ref/ (git-ignored, NOT
gpurun-ignored, so they travel to the GPU box).  No reference source is copied into the
repository.

Usage: python oracle/build_ref.py [--force]
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PIQMC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "piqmc_ref")
MODS = ("sa", "qmc", "tools")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


def ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def built():
    return all(os.path.exists(os.path.join(OUT, m + ext_suffix())) for m in MODS)


def shim(name, src):
    if name in ("sa", "qmc"):
        src = re.sub(r"np\.int_t", "np.int64_t", src)
    if name == "sa":
        src = src.replace("print svec_g.get()", "print(svec_g.get())")
    if name == "tools":
        src = src.replace("J.iterkeys()", "J.keys()")
    return src


def build(force=False):
    if built() and not force:
        return OUT
    if not os.path.isdir(os.path.join(REF, "piqmc")):
        raise RuntimeError("reference tree %s not present and oracle/_ref not prebuilt" % REF)
    import numpy
    from Cython.Build import cythonize  # noqa: F401  (checks Cython is importable)
    tmp = tempfile.mkdtemp(prefix="piqmc_refbuild_")
    try:
        pkg = os.path.join(tmp, "piqmc_ref")
        os.makedirs(pkg)
        open(os.path.join(pkg, "__init__.py"), "w").close()
        for m in MODS:
            with open(os.path.join(REF, "piqmc", m + ".pyx")) as f:
                src = shim(m, f.read())
            with open(os.path.join(pkg, m + ".pyx"), "w") as f:
                f.write(src)
        # cythonize -> C
        subprocess.check_call(
            [sys.executable, "-m", "cython", "-2", "--fast-fail"]
            + [os.path.join("piqmc_ref", m + ".pyx") for m in MODS],
            cwd=tmp)
        inc = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
        os.makedirs(OUT, exist_ok=True)
        for m in MODS:
            so = os.path.join(OUT, m + ext_suffix())
            cmd = [GCC, "-O2", "-fPIC", "-shared", "-fopenmp", "-fno-strict-aliasing",
                   "-Wno-deprecated-declarations", "-Wno-unused-function",
                   "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"] + inc + \
                  [os.path.join(pkg, m + ".c"), "-o", so, "-lm"]
            subprocess.check_call(cmd, cwd=tmp)
        with open(os.path.join(OUT, "__init__.py"), "w") as f:
            f.write("# built by oracle/build_ref.py from the reference's own .pyx files\n")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
