"""The C-ABI shared library: loads without a GPU, exports every symbol that
include/trio_kmers_b200.h declares, keeps the reference's struct prefix, and
fails loudly (no CPU fallback) when no device is present.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "trio_kmers_b200.h")


@pytest.fixture(scope="module")
def kmers():
    from trio_binning_b200 import kmers as mod

    return mod


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", text)
    return sorted(set(names))


def test_header_declares_expected_surface():
    names = declared_functions()
    for legacy in ("create_kmer_hash_set", "count_kmers_in_read", "kmer_to_int", "reverse_complement"):
        assert legacy in names
    assert len(names) >= 25


def test_library_is_named_like_the_reference_extension(kmers):
    # reference kmers.py:30-34 looks for kmers_c<EXT_SUFFIX> next to the module
    base = os.path.basename(kmers._lib_path)
    assert base.startswith("kmers_c.") and base.endswith(".so")
    assert os.path.dirname(kmers._lib_path) == os.path.join(ROOT, "trio_binning_b200")


def test_every_declared_symbol_is_exported(kmers):
    missing = [n for n in declared_functions() if not hasattr(kmers.lib, n)]
    assert missing == []


def test_struct_prefix_matches_reference_layout(kmers):
    # c/kmers.c:12-38 on LP64: kmers@0 full@8 hash_size@16 k@20 num_kmers@24, 32 bytes
    S = kmers._HashSet
    assert (S.kmers.offset, S.full.offset, S.hash_size.offset, S.k.offset, S.num_kmers.offset) == (0, 8, 16, 20, 24)
    assert S.magic.offset >= 28

    class RefStruct(C.Structure):
        _fields_ = [("kmers", C.POINTER(C.c_uint64)), ("full", C.POINTER(C.c_ubyte)),
                    ("hash_size", C.c_int), ("k", C.c_ubyte), ("num_kmers", C.c_int)]
    assert C.sizeof(RefStruct) == 32


def test_host_string_helpers(kmers, golden_kat):
    for s, v in golden_kat["kmer_to_int"]:
        assert kmers.kmer_to_int(s) == v
    for s, v in golden_kat["reverse_complement"]:
        assert kmers.reverse_complement(s) == v
    for x, v in golden_kat["hash_function"]:
        assert kmers.lib.hash_function(C.c_uint64(x)) == v


def test_missing_file_is_ioerror(kmers):
    with pytest.raises(IOError):
        kmers.create_kmer_hash_set("/nonexistent/kmers.txt")


def test_no_gpu_fails_loudly(kmers):
    if kmers.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kmers.create_kmer_hash_set(os.path.join(ROOT, "tests", "golden", "ref_data", "hapA.txt"))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the package may reference it."""
    pkg = os.path.join(ROOT, "trio_binning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text and "_ref/" not in text, f


def test_header_is_plain_c(tmp_path):
    """include/trio_kmers_b200.h is the drop-in boundary: it must compile as C99 (no C++, no CUDA
    types) and declare the reference's symbols with the reference's signatures (c/kmers.c:50,74,185,270)."""
    import shutil
    import subprocess

    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "use_header.c"
    src.write_text(
        '#include "trio_kmers_b200.h"\n'
        "hash_set *(*p_create)(char *) = create_kmer_hash_set;\n"
        "void (*p_count)(char *, hash_set *, hash_set *, int *, int *) = count_kmers_in_read;\n"
        "uint64_t (*p_k2i)(char *, unsigned char) = kmer_to_int;\n"
        "void (*p_rc)(char *, char *, unsigned char) = reverse_complement;\n"
        "char (*p_in)(char *, char *, hash_set *) = kmer_in_hash_set;\n"
        "int f(const tb_stream_options *o) { return o->n_devices + (int)sizeof(tb_packed_batch) + (int)sizeof(tb_stream_stats); }\n")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "use_header.o")], check=True)
