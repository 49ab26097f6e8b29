"""Zero-edit drop-in check, where the reference checkout exists.

The reference's OWN Python modules (copied at test time into a temporary package directory,
never into this repo) are pointed at OUR shared library by giving it the file name their loader
looks for (reference src/trio_binning/kmers.py:30-38), and the reference's OWN, unmodified test
files are run against that package.  Without a GPU only the tests that do not build a k-mer set
can run (string helpers, FASTX I/O, CLI --help): 15 of the reference's 18 tests; the other three
are mirrored in tests/test_gpu_parity.py.  Skipped where /root/reference is absent (GPU box).
"""
import os
import shutil
import subprocess
import sys

import pytest

REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "trio_binning")),
                                reason="reference checkout not present")


@pytest.fixture(scope="module")
def dropin_env(tmp_path_factory):
    from trio_binning_b200 import kmers

    root = tmp_path_factory.mktemp("dropin")
    pkg = root / "trio_binning"
    pkg.mkdir()
    for name in ("kmers.py", "seq.py", "classify_by_kmers.py"):
        shutil.copyfile(os.path.join(REF, "src", "trio_binning", name), pkg / name)
    # our CUDA library under the name the reference's loader scans for
    os.symlink(kmers._lib_path, pkg / os.path.basename(kmers._lib_path))
    env = dict(os.environ, PYTHONPATH=str(root))
    return env


def _run_reference_tests(env, *args):
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", *args]
    return subprocess.run(cmd, env=env, cwd="/tmp", capture_output=True, text=True)


def test_reference_modules_bind_to_our_library(dropin_env):
    code = ("import trio_binning.kmers as k, os; "
            "print(os.path.basename(k.correct_library_file)); "
            "print(k.kmer_to_int('ATGCTAGCTAGAGAGAGAGGA')); print(k.reverse_complement('ACGTN'))")
    p = subprocess.run([sys.executable, "-c", code], env=dropin_env, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.split()
    assert lines[0].startswith("kmers_c.") and lines[1] == "696357446508" and lines[2] == "xACGT"


def test_reference_unit_tests_pass_unmodified(dropin_env):
    p = _run_reference_tests(dropin_env, os.path.join(REF, "tests", "test_kmers.py"),
                             "-k", "kmer_to_int or reverse_complement")
    assert p.returncode == 0, p.stdout + p.stderr
    assert "10 passed" in p.stdout


def test_reference_seq_and_cli_help_tests_pass_unmodified(dropin_env):
    p = _run_reference_tests(dropin_env, os.path.join(REF, "tests", "test_seq.py"),
                             os.path.join(REF, "tests", "test_classify_by_kmers.py"), "-k", "not test_classify_by_kmers or help")
    assert p.returncode == 0, p.stdout + p.stderr
    assert "5 passed" in p.stdout


def test_reference_set_build_fails_loudly_without_gpu(dropin_env):
    """With the reference's own wrapper and no GPU the library reports the failure instead of
    silently computing on the CPU (the reference wrapper then dereferences NULL: we check the
    message our library printed)."""
    from trio_binning_b200 import kmers

    if kmers.device_count() > 0:
        pytest.skip("a GPU is present")
    code = ("import trio_binning.kmers as k; "
            f"s = k.lib.create_kmer_hash_set({os.path.join(REF, 'tests', 'data', 'hapA.txt')!r}.encode()); "
            "print(bool(s))")
    p = subprocess.run([sys.executable, "-c", code], env=dropin_env, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert p.stdout.strip() == "False"
    assert "no CPU fallback" in p.stderr
