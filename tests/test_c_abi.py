"""include/spx.h is a C header: a plain C99 consumer (examples/c_abi_demo.c) compiles against it with gcc -pedantic,
links libspx.so, fails loudly without a device (no CPU fallback) and reproduces the reference's dB line on a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sdr_iq_visualizer_b200", "csrc")


def _build(tmp_path):
    from sdr_iq_visualizer_b200 import _native as nat
    nat.lib()                                     # builds libspx.so when absent
    exe = str(tmp_path / "c_abi_demo")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L" + CSRC, "-lspx", "-lm", "-o", exe],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def _run(exe):
    env = dict(os.environ, LD_LIBRARY_PATH=CSRC + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_c99_and_fails_loudly_without_device(tmp_path):
    from sdr_iq_visualizer_b200 import _native as nat
    exe = _build(tmp_path)
    if nat.device_count() > 0:
        pytest.skip("a GPU is present (covered by the gpu test)")
    res = _run(exe)
    assert res.returncode == 3 and "no CUDA device" in res.stderr


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_consumer_matches_direct_dft(tmp_path):
    res = _run(_build(tmp_path))
    assert res.returncode == 0, res.stdout + res.stderr
    assert "c_abi_demo ok" in res.stdout
