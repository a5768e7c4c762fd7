"""Build the in-tree CUDA library (sm_100a only) with nvcc.

`python -m surfd_b200.build` compiles surfd_b200/csrc/*.cu into surfd_b200/_surfd_b200.so.  The .so is
git-ignored but travels to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
import os, subprocess, sys, hashlib, shutil

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_surfd_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", 
          "-DSURFD_BUILDING", "--expt-relaxed-constexpr"]
# per-file extra flags
EXTRA = {
    # the replay must match the reference's unfused IEEE arithmetic; SURFD_MC_FLAGS=-DMC_PROFILE adds cycle counters
    "mc.cu": ["-fmad=false"] + os.environ.get("SURFD_MC_FLAGS", "").split(),
}


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(path, flags):
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    for dep in sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "surfd_b200.h")]:
        p = os.path.join(CSRC, dep)
        if os.path.isfile(p) and (dep.endswith((".cu", ".cuh", ".h", ".inc"))):
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def build(verbose=False, force=False, variant=None, extra_flags=None):
    """variant / extra_flags: a diagnostic build of the same sources next to the product (e.g. variant="mcprof",
    extra_flags={"mc.cu": ["-DMC_PROFILE"]} -> surfd_b200/_surfd_b200_mcprof.so; select it with SURFD_B200_LIB)."""
    nvcc = _nvcc()
    global OBJ, OUT
    if variant:
        OBJ = os.path.join(HERE, "csrc", "_obj_" + variant)
        OUT = os.path.join(HERE, "_surfd_b200_%s.so" % variant)
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    changed = False
    for src in sources():
        flags = ARCH + COMMON + EXTRA.get(src, []) + (extra_flags or {}).get(src, [])
        obj = os.path.join(OBJ, src[:-3] + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp(os.path.join(CSRC, src), flags)
        old = open(stamp_file).read() if os.path.exists(stamp_file) else ""
        if force or old != stamp or not os.path.exists(obj):
            cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            with open(stamp_file, "w") as f:
                f.write(stamp)
            changed = True
        objs.append(obj)
    if changed or not os.path.exists(OUT):
        cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    if "--mcprof" in sys.argv:
        print(build(verbose=True, variant="mcprof", extra_flags={"mc.cu": ["-DMC_PROFILE"]}))
    else:
        print(build(verbose=True, force="--force" in sys.argv))
