"""The C-ABI is a plain C boundary: include/boojum_b200.h must compile as strict C99 and a C program (examples/ntt_host.c, no
Python / torch involved) must link against the library.  Without a GPU the program has to stop at bj_ctx_create with
BJ_ERR_NO_DEVICE (no CPU fallback); on a GPU box it must complete the NTT round trip."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")


def _build(tmp_path):
    exe = str(tmp_path / "ntt_host")
    lib_dir = os.path.join(ROOT, "era_boojum_b200")
    subprocess.check_call([GCC, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "ntt_host.c"), "-L", lib_dir, "-lboojum_b200", "-Wl,-rpath," + lib_dir, "-o", exe])
    return exe


@pytest.mark.skipif(GCC is None, reason="no C compiler")
def test_header_is_c99_and_client_links_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "round trip ok" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr and "-3" in r.stderr


@pytest.mark.gpu
@pytest.mark.skipif(GCC is None, reason="no C compiler")
def test_c_client_round_trip_on_gpu(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "round trip ok" in r.stdout and "kernel launches" in r.stdout
