"""Runs the reference's own test files UNCHANGED against pytorch_hmm_b200.

    python tools/run_ref_suite.py --vendor     # build container: copies /root/reference/tests/test_*.py into ref_suite/_vendored/
                                               # (git-ignored -- reference sources are never committed -- but it travels with gpurun)
    python tools/run_ref_suite.py --run        # GPU box: pytest over the vendored files, `import pytorch_hmm` aliased by ref_suite/alias.py
    python tools/run_ref_suite.py --clean      # remove the vendored copies again (do this after the run: they are reference sources)

Typical use from the build container:
    python tools/run_ref_suite.py --vendor && gpurun -- 'python tools/run_ref_suite.py --run > gpurun_out/ref_suite.log 2>&1' ; python tools/run_ref_suite.py --clean
The last result is kept in profiles/r02_reference_suite.txt.
"""
import argparse
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDOR = os.path.join(ROOT, "ref_suite", "_vendored")
FILES = ["test_hmm.py", "test_mixture_gaussian.py", "test_hsmm.py", "test_streaming.py"]
CONFTEST = '''import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ref_suite.alias import install
install()
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vendor", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--clean", action="store_true")
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    if a.vendor:
        os.makedirs(VENDOR, exist_ok=True)
        for f in FILES:
            shutil.copy(os.path.join(a.ref, "tests", f), os.path.join(VENDOR, f))
        with open(os.path.join(VENDOR, "conftest.py"), "w") as fh:
            fh.write(CONFTEST)
        print("vendored", FILES, "->", VENDOR)
    if a.clean:
        shutil.rmtree(VENDOR, ignore_errors=True)
        print("removed", VENDOR)
    if a.run:
        if not os.path.isdir(VENDOR):
            raise SystemExit("ref_suite/_vendored is missing: run with --vendor in the build container first")
        cmd = [sys.executable, "-m", "pytest", VENDOR, "-q", "-p", "no:cacheprovider", "-o", "addopts=", "-rs", "--tb=line"]
        raise SystemExit(subprocess.call(cmd, cwd=ROOT))


if __name__ == "__main__":
    main()
