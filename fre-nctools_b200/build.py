"""Build libxgrid_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libxgrid_b200.so")
CLI_SRC = os.path.join(HERE, "cli", "fregrid_b200.c")
CLI_OUT = os.path.join(HERE, "bin", "fregrid_b200")
# -fmad=false: the exchange grid's accept/reject predicates must round like the reference's
# un-fused x86-64 arithmetic (see csrc/xgrid_geom.cuh).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--shared", "-Xptxas", "-v",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") or f.endswith(".c"))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", "xgrid_b200.h"), os.path.abspath(__file__)]
    deps += [os.path.join(HERE, "cli", f) for f in os.listdir(os.path.join(HERE, "cli"))]
    if not all(os.path.exists(os.path.join(HERE, "bin", t)) for t in ("fregrid_b200", "make_coupler_mosaic_b200")):
        return True
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: developer builds of kernel variants next to the product library (scripts/clip_variants.py)."""
    if out is None and not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out or OUT] + [os.path.join(CSRC, s) for s in sources()] + ["-lcudart", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libxgrid_b200.so")
    if out is not None:
        return out
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(r.stdout + r.stderr)
    build_cli()
    return OUT


CLI_TOOLS = ("fregrid_b200", "make_coupler_mosaic_b200")


def build_cli():
    """fregrid_b200, make_coupler_mosaic_b200: the host side of the reference command lines in C, linked against
    libxgrid_b200.so next to them."""
    os.makedirs(os.path.dirname(CLI_OUT), exist_ok=True)
    for tool in CLI_TOOLS:
        src = os.path.join(HERE, "cli", tool + ".c")
        out = os.path.join(HERE, "bin", tool)
        cmd = [os.environ.get("CC", "gcc"), "-O2", "-std=gnu99", "-Wall", "-o", out, src, "-L" + HERE, "-lxgrid_b200",
               "-Wl,-rpath,$ORIGIN/..", "-Wl,--allow-shlib-undefined", "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed building " + tool)
    return CLI_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
