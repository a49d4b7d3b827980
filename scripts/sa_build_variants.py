"""Builds experiment variants of the CUDA library HERE (no GPU needed) so that a GPU call only measures: every argument
`name=-Dflag -Dflag` recompiles csrc/sparse_align.cu (or the source named by --src=) with those switches and links it with the
default build's other objects into dsdtm_b200/lib_exp/<name>/libdsdtm_gpu.so (git-ignored, travels with gpurun).
Use: python scripts/sa_build_variants.py base="-DDSDTM_SA_CVT=1 -DDSDTM_SA_STAGE=0" cvt3="-DDSDTM_SA_CVT=3 -DDSDTM_SA_STAGE=0"
then on the box: DSDTM_GPU_LIB=dsdtm_b200/lib_exp/cvt3/libdsdtm_gpu.so python scripts/sa_sweep.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dsdtm_b200 import build as B


def main():
    src = "sparse_align.cu"
    args = [a for a in sys.argv[1:]]
    if args and args[0].startswith("--src="):
        src = args.pop(0)[6:]
    B.build()
    for a in args:
        name, flags = a.split("=", 1)
        d = os.path.join(B.HERE, "lib_exp", name)
        os.makedirs(d, exist_ok=True)
        o = os.path.join(d, src.replace(".cu", ".o"))
        r = subprocess.run([B.nvcc()] + B.NVCC_FLAGS + flags.split() + ["-c", os.path.join(B.CSRC, src), "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr); raise SystemExit(1)
        objs = [o if s == src else os.path.join(B.LIBDIR, s.replace(".cu", ".o")) for s in B.SOURCES]
        subprocess.run([B.nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(d, "libdsdtm_gpu.so")] + objs, check=True)
        regs = [l for l in r.stderr.splitlines() if "Used" in l]
        print(name, flags, "->", os.path.join(d, "libdsdtm_gpu.so"), "|", len(regs), "kernels")


if __name__ == "__main__":
    main()
