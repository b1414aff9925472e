"""Compiles the unmodified C++ shim (integration/src/*.cpp) + tests/cpp/test_shim.cpp with g++ against integration/stubs
(minimal OpenCV / Eigen look-alikes; the image has neither) and links it to libsivo_b200.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "build", "test_shim")
SOURCES = [os.path.join(ROOT, "integration", "src", "bayesian_segnet.cpp"), os.path.join(ROOT, "integration", "src", "ORBextractor.cc"),
           os.path.join(ROOT, "tests", "cpp", "test_shim.cpp")]


def build(force=False):
    deps = SOURCES + [os.path.join(ROOT, "sivo_b200", "libsivo_b200.so"), os.path.join(ROOT, "include", "sivo_b200.h")]
    for d, _, files in os.walk(os.path.join(ROOT, "integration")):
        deps += [os.path.join(d, f) for f in files]
    if not force and os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    lib_dir = os.path.join(ROOT, "sivo_b200")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "integration", "stubs"),
           "-I" + os.path.join(ROOT, "integration", "include"), "-I" + os.path.join(ROOT, "include"), *SOURCES,
           "-L" + lib_dir, "-lsivo_b200", "-Wl,-rpath," + lib_dir, "-o", BIN]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("shim build failed:\n" + r.stderr[-4000:])
    return BIN
