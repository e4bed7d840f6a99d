#!/usr/bin/env python3
"""Build the reference's UNCHANGED Cython wrapper (pspace/PSPACE.pyx + PSPACE.pxd, read where they lie under
/root/reference) against THIS repo's headers (pspace_b200/host/include) and libpspace.so — the drop-in check.

Outputs (build artefacts only, git-ignored, shipped to the GPU box with the snapshot):
    oracle/_ref/dropin/real/pspace/PSPACE.<abi>.so        linked to pspace_b200/lib/libpspace.so
    oracle/_ref/dropin/complex/pspace/PSPACE.<abi>.so     linked to pspace_b200/lib/complex/libpspace.so (-DUSE_COMPLEX)
plus, beside each, an __init__.py and a _pointer_utils.py written HERE (three lines each; the compiled wrapper
imports `pspace._pointer_utils.ensure_scalar_pointer / ensure_int_pointer` at run time) — no reference source
file is copied into the repo; the .pyx/.pxd are cythonized from a scratch copy under /tmp.
A second pair, oracle/_ref/dropin_additive/{real,complex}, is the same wrapper with the ADDITIVE binding of
integration/PSPACE_additive.{pxd,pyx}.inc applied to the scratch copy (five extern lines inserted into the cppclass
block of PSPACE.pxd, four def methods appended to PyParameterContainer in PSPACE.pyx — what INTEGRATION.md shows as a
patch): projectVector / projectMatrix / evaluateExpansion / evalBasis reach the GPU from the reference's own wrapper.
Runs only where /root/reference exists (the build container).  tests/test_gpu_dropin.py loads the result.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PSPACE_REFERENCE", "/root/reference")

POINTER_UTILS = '''"""Run-time helper the compiled reference wrapper imports (written for the drop-in check)."""
import numpy as np


def ensure_scalar_pointer(values, dtype=np.double):
    a = np.ascontiguousarray(values, dtype=dtype)
    return a, a.ctypes.data


def ensure_int_pointer(values, dtype=np.intc):
    a = np.ascontiguousarray(values, dtype=dtype)
    return a, a.ctypes.data
'''


PXD_ANCHOR = "        void initializeQuadrature(const int *nqpts)\n"


def apply_additive(pkg):
    """The additive binding: insert the extern lines after the last method of `cdef cppclass ParameterContainer`, append the
    def methods to `cdef class PyParameterContainer` (the last class of PSPACE.pyx)."""
    inc = os.path.join(ROOT, "integration")
    pxd = open(os.path.join(pkg, "PSPACE.pxd")).read()
    assert pxd.count(PXD_ANCHOR) == 1, "PSPACE.pxd changed: anchor for the additive externs not found"
    pxd = pxd.replace(PXD_ANCHOR, PXD_ANCHOR + open(os.path.join(inc, "PSPACE_additive.pxd.inc")).read())
    open(os.path.join(pkg, "PSPACE.pxd"), "w").write(pxd)
    pyx = open(os.path.join(pkg, "PSPACE.pyx")).read()
    assert pyx.rstrip().endswith("return"), "PSPACE.pyx changed: PyParameterContainer is expected to be the last class"
    open(os.path.join(pkg, "PSPACE.pyx"), "w").write(pyx.rstrip("\n") + "\n" + open(os.path.join(inc, "PSPACE_additive.pyx.inc")).read())


def build(complex_, additive=False):
    import numpy
    from Cython.Build import cythonize
    tag = "complex" if complex_ else "real"
    out_pkg = os.path.join(ROOT, "oracle", "_ref", "dropin_additive" if additive else "dropin", tag, "pspace")
    os.makedirs(out_pkg, exist_ok=True)
    libdir = os.path.join(ROOT, "pspace_b200", "lib", "complex" if complex_ else "")
    libdir = os.path.normpath(libdir)
    with tempfile.TemporaryDirectory(prefix="pspace_dropin_") as tmp:
        pkg = os.path.join(tmp, "pspace")
        os.makedirs(pkg)
        for f in ("PSPACE.pyx", "PSPACE.pxd"):
            shutil.copy(os.path.join(REF, "pspace", f), pkg)          # scratch copy, never enters the repo
        if additive:
            apply_additive(pkg)
        open(os.path.join(pkg, "__init__.py"), "w").close()
        with open(os.path.join(pkg, "_config.pxi"), "w") as fh:
            fh.write(f"DEF USE_COMPLEX = {1 if complex_ else 0}\n")
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            from setuptools import Extension
            ext = Extension("pspace.PSPACE", sources=["pspace/PSPACE.pyx"], language="c++",
                            include_dirs=[numpy.get_include(), pkg, os.path.join(ROOT, "pspace_b200", "host", "include")],
                            libraries=["pspace"], library_dirs=[libdir], runtime_library_dirs=[libdir],
                            define_macros=[("USE_COMPLEX", None)] if complex_ else [])
            exts = cythonize([ext], include_path=[pkg, tmp], quiet=True, force=True,
                             compiler_directives={"embedsignature": True, "binding": True})
            cpp = exts[0].sources[0]
            so = os.path.join(out_pkg, "PSPACE" + sysconfig.get_config_var("EXT_SUFFIX"))
            cmd = ["g++", "-shared", "-fPIC", "-O2", "-w", "-std=c++14", cpp, "-o", so,
                   "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), "-I" + pkg,
                   "-I" + os.path.join(ROOT, "pspace_b200", "host", "include"),
                   "-L" + libdir, "-lpspace", "-Wl,-rpath," + libdir]
            if complex_:
                cmd.insert(1, "-DUSE_COMPLEX")
            subprocess.check_call(cmd)
        finally:
            os.chdir(cwd)
    with open(os.path.join(out_pkg, "_pointer_utils.py"), "w") as fh:
        fh.write(POINTER_UTILS)
    with open(os.path.join(out_pkg, "__init__.py"), "w") as fh:
        fh.write('"""drop-in check package: the reference wrapper compiled against pspace_b200\'s libpspace.so"""\n')
    print("built", so)


def main():
    if not os.path.isdir(os.path.join(REF, "pspace")):
        print("reference checkout not present; keeping prebuilt oracle/_ref/dropin (if any)")
        return 0
    build(False)
    build(True)
    build(False, additive=True)
    build(True, additive=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
