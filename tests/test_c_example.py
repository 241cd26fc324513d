"""The C ABI is usable from plain C: the header compiles as C99 (CPU check), and examples/siglip_c_demo.c — no Python,
no torch in the process — runs the fused step and checks it against a double-precision host evaluation (GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "distributed_sigmoid_loss_b200")


def test_header_is_valid_c99(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    src = tmp_path / "hdr.c"
    src.write_text('#include "siglip_b200.h"\nint main(void) { return siglip_ctx_handle_bytes() == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_plain_c_program_runs_the_fused_step(tmp_path):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = str(tmp_path / "siglip_c_demo")
    r = subprocess.run(["nvcc", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "examples", "siglip_c_demo.c"),
                        "-I", os.path.join(ROOT, "include"), "-L", LIBDIR, "-lsiglip_b200",
                        "-Xlinker", "-rpath", "-Xlinker", LIBDIR], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C-ABI DEMO PASS" in r.stdout, r.stdout + r.stderr
