"""Builds the C++ host adapters (reference class interfaces over the C-ABI). Filled in by dsdtm_b200/host/*.cpp."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "lib", "libdsdtm_host.so")


def sources():
    return sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cpp"))


def build(force=False):
    src = sources()
    if not src:
        return None
    deps = src + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".h")] + \
        [os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "dsdtm_gpu.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    gpu = os.path.join(os.path.dirname(LIB), "libdsdtm_gpu.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-o", LIB] + src + \
          ["-I", os.path.join(os.path.dirname(os.path.dirname(HERE)), "include"), gpu, "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return LIB
