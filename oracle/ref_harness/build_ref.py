"""Harness (this container only): compile the reference's two Cython modules
from where they lie under /root/reference into oracle/_ref/cy_src/*.so.

The sources are copied to a scratch dir under /tmp (the reference tree is
read-only and ships stale cpython-310 binaries), one line is patched there
(cy_src/lgmres_cy.pyx:510 returns a memoryview where :331 returns the ndarray,
which makes every inequality config raise TypeError), and only the built
shared objects land in oracle/_ref/ (git-ignored).  Nothing from /root/reference
is copied into the repository.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

REF = os.environ.get("TTIPM_REF_TREE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.abspath(os.path.join(HERE, "..", "_ref"))

SETUP = r'''
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy as np
kw = dict(include_dirs=[np.get_include()],
          define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
          extra_compile_args=["-O2"])
setup(ext_modules=cythonize([
        Extension("cy_src.tt_ops_cy", ["cy_src/tt_ops_cy.pyx"], **kw),
        Extension("cy_src.lgmres_cy", ["cy_src/lgmres_cy.pyx"], **kw)],
      language_level=3, force=True),
      script_args=["build_ext", "--inplace"])
'''


def build(force=False):
    dst = os.path.join(OUT, "cy_src")
    if not force and len(glob.glob(os.path.join(dst, "*.so"))) == 2:
        return True
    if not os.path.isdir(os.path.join(REF, "cy_src")):
        return False
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="ttipm_ref_") as tmp:
        os.makedirs(os.path.join(tmp, "cy_src"))
        for f in ("tt_ops_cy.pyx", "lgmres_cy.pyx"):
            shutil.copy(os.path.join(REF, "cy_src", f), os.path.join(tmp, "cy_src", f))
            os.chmod(os.path.join(tmp, "cy_src", f), 0o644)
        p = os.path.join(tmp, "cy_src", "lgmres_cy.pyx")
        lines = open(p).read().split("\n")
        assert lines[509].strip() == "return self.flat_result", lines[509]
        lines[509] = lines[509].replace("self.flat_result", "self.flat_result_arr")
        open(p, "w").write("\n".join(lines))
        open(os.path.join(tmp, "cy_src", "__init__.py"), "w").write("")
        open(os.path.join(tmp, "setup.py"), "w").write(SETUP)
        env = dict(os.environ)
        env["PYTHONPATH"] = os.path.join(HERE, "standins") + os.pathsep + env.get("PYTHONPATH", "")
        subprocess.check_call([sys.executable, "setup.py"], cwd=tmp, env=env,
                              stdout=subprocess.DEVNULL)
        for so in glob.glob(os.path.join(tmp, "cy_src", "*.so")):
            shutil.copy(so, dst)
    open(os.path.join(dst, "__init__.py"), "w").write("")
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("reference built into", OUT if ok else "(reference absent)")
