"""Host logic (restart schedules, the batched restart driver, budgets, printing) against a CPU stand-in of the C ABI
(tests/mock/bbgpu_mock.cpp), under AddressSanitizer + UBSan.  No GPU, no optimisation: control flow only."""
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_classes_under_sanitizers(tmp_path):
    exe = str(tmp_path / "host_driver")
    srcs = [os.path.join(ROOT, "tests", "mock", "bbgpu_mock.cpp"), os.path.join(ROOT, "tests", "mock", "host_driver.cpp")]
    srcs += sorted(glob.glob(os.path.join(ROOT, "bboptpy_b200", "host", "*.cpp")))
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-std=c++17",
                    "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "bboptpy_b200", "host")] + srcs + ["-o", exe],
                   check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "batch kind 0: 37 runs" in r.stdout and "batch kind 1: 37 runs" in r.stdout
