"""CPU check of the ordered-sum arithmetic core (patolette_b200/csrc/pb_span.h): the header the CUDA
kernels use is compiled for the host and run against the literal sequential loop on adversarial data
(ties, sums hovering around zero / around powers of two, cancellation, wide dynamic range)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_span_arithmetic_matches_sequential_loop(tmp_path):
    exe = str(tmp_path / "test_span")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-I", os.path.join(ROOT, "patolette_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "test_span.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK")
